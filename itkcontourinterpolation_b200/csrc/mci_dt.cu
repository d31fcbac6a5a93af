// mci_dt.cu -- the Interpolate1to1 recursion with the distance-transform median
// (reference X:556-793 with FindMedianImageDistances X:405-516).
//
// What runs (round 2):
//   k_dt_tree          ONE persistent launch: a CTA takes a node from a global ready queue and runs it from start to
//                      end (prep, d2 histograms, scan, median, mid shape), then publishes its children; rows and
//                      histograms live in shared memory.  See the comment above the kernel.
//   k_compose_*        volume writes of all mid slices perpendicular to z, then y: shapes binned by plane, one warp per
//                      row, plain loads / stores (X:743-764 with X:1623-1636 folded in)
//   k_apply_axis0      volume writes of the trees perpendicular to x, x-run by x-run
//   k_apply_masks      replay of another GPU's mid slices (segment-sharded multi-GPU combine)
//
// Kept for A/B runs (MCI_DT_LEVELS=1), the round-1 form of the same recursion as six grid-wide kernels per level:
//   k_dt_prep    per node : translation halving, union region, translate+binarise+AND into bit rows,
//                           per-row extents of the intersection, work items            (X:570-666)
//   k_dt_hist    per item : exact squared distance d2 of every (i xor j) pixel to the intersection,
//                           histograms per side over bin = PixelType(10 * sqrtf(d2))     (X:413-459)
//   k_dt_scan    per node : prefix sums, first bin minimising the score, threshold       (X:461-494)
//   k_dt_median  per item : median = intersection | (xor pixels with d2 <= threshold), bounding box (X:497-515)
//   k_dt_alloc   per node : crop/clip, allocate the mid shape, spawn the children        (X:679-729, 766-792)
//   k_dt_emit    per item : store the mid shape, write it into the volume with the max rule (X:743-764)
// (work items there are fixed-size word ranges of a node's bit rows, so the whole GPU works on a level no matter how
// few nodes it has or how unequal they are; the price is 18 dependent launches for a gap of 8).
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "mci_kernels.cuh"
#include "mci_launch.h"

namespace mci
{

#define DT_CHUNK 256 // words of a node's bit rows per work item (one CTA: 8 warps x 32 words)
#define DT_HCHUNK MCI_DT_HCHUNK // words per histogram work item (one warp), emitted only where (i xor j) pixels exist
#define DT_BIG_STRIDE (2 * 65536 + 2048) // wrapped-bin table: counts of both sides + bitmap of touched bins

__device__ __forceinline__ uint32_t
pack_row(int a, int b, bool single)
{
  return (uint32_t)a | ((uint32_t)b << 13) | (single ? (1u << 26) : 0u) | (1u << 27);
}

// exact squared distance from pixel (x, r) to the nearest intersection pixel, or `cap` if not below cap.
// rowInfo[r] packs the first / last set column of row r and whether the row is one run (then the
// distance to the row is the distance to that interval); 0 = empty row.  bandInfo[b] is the same
// extent over rows 8b..8b+7: a whole band is skipped when even its nearest row and nearest column
// cannot beat the best distance so far.
// NC = true: the rows are read through the non-coherent path (__ldg; written by an earlier kernel); NC = false:
// plain loads (k_dt_tree writes and reads them in one kernel, in shared memory when they fit)
template <bool NC>
__device__ __forceinline__ uint32_t
dt_ld(const uint32_t * p)
{
  return NC ? __ldg(p) : *p;
}

template <bool NC = true>
__device__ __forceinline__ int
dt_probe_row(const uint32_t * Xb, const uint32_t * rowInfo, int pitch, int x, int rr, int dd, int best)
{
  const uint32_t info = dt_ld<NC>(rowInfo + rr);
  if (!info)
    return best;
  const int a = info & 0x1FFF, b = (info >> 13) & 0x1FFF;
  const int dx = x < a ? a - x : (x > b ? x - b : 0);
  const int lo = dd + dx * dx;
  if (lo >= best)
    return best;
  if ((info >> 26) & 1u)
    return lo;
  return min(best, dd + row_nearest_sq(Xb + (size_t)rr * pitch, pitch, x, best - dd));
}

template <bool NC = true>
__device__ __forceinline__ bool
dt_band_prunable(const uint32_t * bandInfo, int band, int x, int dd, int best)
{
  const uint32_t info = dt_ld<NC>(bandInfo + band);
  if (!info)
    return true;
  const int a = info & 0x1FFF, b = (info >> 13) & 0x1FFF;
  const int dx = x < a ? a - x : (x > b ? x - b : 0);
  return dd + dx * dx >= best;
}

template <bool NC = true>
__device__ __forceinline__ int
dt_edt_sq(const uint32_t * Xb, const uint32_t * rowInfo, int pitch, int h, int xr0, int xr1, int x, int r, int cap)
{
  const uint32_t * bandInfo = rowInfo + h;
  int best = cap;
  // rows at and below r
  for (int rr = max(r, xr0); rr <= xr1;)
  {
    const int dy = rr - r, dd = dy * dy;
    if (dd >= best)
      break;
    if ((rr & 7) == 0 && dt_band_prunable<NC>(bandInfo, rr >> 3, x, dd, best))
    {
      rr += 8;
      continue;
    }
    best = dt_probe_row<NC>(Xb, rowInfo, pitch, x, rr, dd, best);
    ++rr;
  }
  // rows above r
  for (int rr = min(r - 1, xr1); rr >= xr0;)
  {
    const int dy = r - rr, dd = dy * dy;
    if (dd >= best)
      break;
    if ((rr & 7) == 7 && dt_band_prunable<NC>(bandInfo, rr >> 3, x, dd, best))
    {
      rr -= 8;
      continue;
    }
    best = dt_probe_row<NC>(Xb, rowInfo, pitch, x, rr, dd, best);
    --rr;
  }
  return best;
}

// Histogram bin of a pixel at squared distance d2 from the intersection.  The reference's sdf is ITK's signed
// Maurer map with SquaredDistance off (its default; X:395-402 never sets it), i.e. float(sqrt(d2)) outside the
// object, and `PixelType dist = fractioning * sdf` (X:431) is a float multiply, a truncating conversion and a
// narrowing to the pixel type.  __fsqrt_rn / __fmul_rn are the same IEEE operations the host code performs.
template <typename T>
__device__ __forceinline__ int
dt_bin(int d2, int * err)
{
  int q = __float2int_rz(__fmul_rn(10.0f, __fsqrt_rn((float)d2)));
  if (sizeof(T) == 1)
    return q & 0xFF;
  if (std::is_signed<T>::value)
  {
    short v = (short)q;
    if (v < 0)
    {
      atomicMax(err, ERR_PIXELTYPE); // the reference indexes its histogram with a negative value here
      v = 0;
    }
    return v;
  }
  return q & 0xFFFF;
}

// `sdf <= float(bestBin) / fractioning` (X:497-508) as a threshold on the integer d2: the largest n with
// float(sqrt(n)) <= upper (sqrt is monotone, so {sdf <= upper} = {d2 <= n}).
__device__ __forceinline__ int
dt_threshold_d2(int bestBin)
{
  const float upper = __fdiv_rn((float)bestBin, 10.0f);
  int n = (int)(upper * upper);
  while (__fsqrt_rn((float)(n + 1)) <= upper)
    ++n;
  while (n > 0 && __fsqrt_rn((float)n) > upper)
    --n;
  return n;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dt_prep(const Node * __restrict__ nodes, const int * __restrict__ count, NodeWork * work, const uint32_t * __restrict__ arena,
          uint32_t * scratch, unsigned long long * scratchTop, unsigned long long scratchCap, DtItem * items, int * itemCount,
          int itemCap, DtItem * hitems, int * hitemCount, int hitemCap, int * err)
{
  pdl_wait();
  __shared__ int s_r0, s_r1;
  __shared__ unsigned long long s_off;
  __shared__ int s_item;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = *count;
  for (int nidx = blockIdx.x; nidx < n; nidx += gridDim.x)
  {
    const Node nd = nodes[nidx];
    // translation halving with carry (X:570-608)
    int t[2] = { nd.tx, nd.ty }, iT[2], jT[2];
    bool carry = false;
#pragma unroll
    for (int d = 0; d < 2; ++d)
    {
      if (!carry)
      {
        iT[d] = t[d] / 2;
        carry = (t[d] % 2) != 0;
      }
      else if (t[d] % 2 == 0)
        iT[d] = t[d] / 2;
      else
      {
        iT[d] = t[d] > 0 ? t[d] / 2 + 1 : t[d] / 2 - 1;
        carry = false;
      }
      jT[d] = iT[d] - t[d];
    }
    const int mid = (nd.i + nd.j + (carry ? 1 : 0)) / 2;
    const int ux0 = min(nd.A.x0 + iT[0], nd.B.x0 + jT[0]), uy0 = min(nd.A.y0 + iT[1], nd.B.y0 + jT[1]);
    const int ux1 = max(nd.A.x0 + nd.A.w + iT[0], nd.B.x0 + nd.B.w + jT[0]);
    const int uy1 = max(nd.A.y0 + nd.A.h + iT[1], nd.B.y0 + nd.B.h + jT[1]);
    const int w = ux1 - ux0, h = uy1 - uy0;
    const int pitch = (w + 31) >> 5;
    const int words = h * pitch;
    __syncthreads();
    if (tid == 0)
    {
      s_r0 = 0x7FFFFFFF;
      s_r1 = -1;
      s_off = atomicAdd(scratchTop, (unsigned long long)(3 * (size_t)words + h + (h + 7) / 8));
    }
    __syncthreads();
    const unsigned long long off = s_off;
    NodeWork * W = work + nidx;
    if (off + 3ull * words + h + (h + 7) / 8 > scratchCap || w > 8191)
    {
      if (tid == 0)
      {
        atomicMax(err, ERR_SCRATCH);
        W->alive = 0;
      }
      continue;
    }
    uint32_t * Xb = scratch + off;
    uint32_t * Ib = Xb + words;
    uint32_t * Sb = Ib + words;
    uint32_t * rowInfo = Sb + words;
    // translate, binarise, intersect (X:610-666); a group of GW lanes per row (GW = row pitch rounded up to a power of
    // two, 32 / GW rows per warp at once) so that row extents come for free and narrow shapes still fill the warp
    const int GW = pitch <= 1 ? 1 : (pitch <= 2 ? 2 : (pitch <= 4 ? 4 : (pitch <= 8 ? 8 : (pitch <= 16 ? 16 : 32))));
    const int RPW = 32 / GW, sub = lane / GW, sl = lane - sub * GW;
    int r0 = 0x7FFFFFFF, r1 = -1;
    for (int rb = warp * RPW; rb < h; rb += 8 * RPW)
    {
      const int r = rb + sub;
      const bool rowOk = r < h;
      const int y = uy0 + r;
      int first = 0x7FFFFFFF, last = -1, runs = 0;
      uint32_t prevTop = 0;
      for (int wb = 0; wb < pitch; wb += GW)
      {
        const int wi = wb + sl;
        uint32_t a = 0, b = 0;
        if (rowOk && wi < pitch)
        {
          const int x = ux0 + 32 * wi;
          a = mask_word(arena, nd.A, y - iT[1], x - iT[0]);
          b = mask_word(arena, nd.B, y - jT[1], x - jT[0]);
          const size_t idx = (size_t)r * pitch + wi;
          Xb[idx] = a & b;
          Ib[idx] = a;
          Sb[idx] = a ^ b;
        }
        const uint32_t xw = a & b;
        // run starts in this word: set bits whose lower neighbour (previous bit, or top bit of the previous word) is clear
        uint32_t below = __shfl_up_sync(0xFFFFFFFFu, xw, 1, GW);
        if (sl == 0)
          below = prevTop;
        const uint32_t starts = xw & ~((xw << 1) | (below >> 31));
        runs += __popc(starts);
        if (xw)
        {
          first = min(first, 32 * wi + __ffs(xw) - 1);
          last = max(last, 32 * wi + 31 - __clz(xw));
        }
        prevTop = __shfl_sync(0xFFFFFFFFu, xw, GW - 1, GW);
      }
      for (int o = GW >> 1; o > 0; o >>= 1)
      {
        first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
        last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
        runs += __shfl_xor_sync(0xFFFFFFFFu, runs, o);
      }
      if (sl == 0 && rowOk)
      {
        rowInfo[r] = last >= 0 ? pack_row(first, last, runs == 1) : 0u;
        if (last >= 0)
        {
          r0 = min(r0, r);
          r1 = max(r1, r);
        }
      }
    }
    // rows of the whole warp (one leader lane per group held them)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      r0 = min(r0, __shfl_xor_sync(0xFFFFFFFFu, r0, o));
      r1 = max(r1, __shfl_xor_sync(0xFFFFFFFFu, r1, o));
    }
    if (lane == 0 && r1 >= 0)
    {
      atomicMin(&s_r0, r0);
      atomicMax(&s_r1, r1);
    }
    __syncthreads();
    // extents of 8-row bands
    for (int bnd = tid; bnd < (h + 7) / 8; bnd += 256)
    {
      int a = 0x7FFFFFFF, b = -1;
      for (int q = 0; q < 8 && bnd * 8 + q < h; ++q)
      {
        const uint32_t info = rowInfo[bnd * 8 + q];
        if (info)
        {
          a = min(a, (int)(info & 0x1FFF));
          b = max(b, (int)((info >> 13) & 0x1FFF));
        }
      }
      rowInfo[h + bnd] = b >= 0 ? pack_row(a, b, false) : 0u;
    }
    // histogram work items: DT_HCHUNK-word chunks that hold (i xor j) pixels (skipped when the intersection is empty)
    if (s_r1 >= 0)
    {
      const int nChunks = (words + DT_HCHUNK - 1) / DT_HCHUNK;
      for (int cb = 0; cb < nChunks; cb += 256)
      {
        const int c = cb + tid;
        bool any = false;
        if (c < nChunks)
          for (int q = 0; q < DT_HCHUNK && c * DT_HCHUNK + q < words; ++q)
            any |= Sb[c * DT_HCHUNK + q] != 0u;
        const unsigned vote = __ballot_sync(0xFFFFFFFFu, any);
        int base = 0;
        if (lane == 0 && vote)
          base = atomicAdd(hitemCount, __popc(vote));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (any)
        {
          const int pos = base + __popc(vote & ((1u << lane) - 1u));
          if (pos < hitemCap)
          {
            DtItem it;
            it.node = nidx;
            it.w0 = c * DT_HCHUNK;
            hitems[pos] = it;
          }
          else
            atomicMax(err, ERR_SCRATCH);
        }
      }
    }
    if (tid == 0)
    {
      NodeWork nw;
      nw.nd = nd;
      nw.ux0 = ux0;
      nw.uy0 = uy0;
      nw.w = w;
      nw.h = h;
      nw.pitch = pitch;
      nw.words = words;
      nw.iTx = iT[0];
      nw.iTy = iT[1];
      nw.jTx = jT[0];
      nw.jTy = jT[1];
      nw.mid = mid;
      nw.xr0 = s_r0;
      nw.xr1 = s_r1;
      nw.bigIdx = -1;
      nw.thr = 0;
      nw.bx0 = nw.by0 = 0x7FFFFFFF;
      nw.bx1 = nw.by1 = -0x7FFFFFFF;
      nw.scratchOff = off;
      nw.alive = s_r1 >= 0 ? 1 : 0; // empty intersection: both median finders give an empty image
      *W = nw;
      if (nw.alive)
      {
        const int nItems = (words + DT_CHUNK - 1) / DT_CHUNK;
        int base = atomicAdd(itemCount, nItems);
        if (base + nItems > itemCap)
        {
          atomicMax(err, ERR_SCRATCH);
          W->alive = 0;
        }
        else
          for (int q = 0; q < nItems; ++q)
          {
            DtItem it;
            it.node = nidx;
            it.w0 = q * DT_CHUNK;
            items[base + q] = it;
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int
dt_get_big(NodeWork * W, int * bigCount, int nbig, int * err)
{
  int bi = atomicAdd(&W->bigIdx, 0);
  if (bi >= 0)
    return bi;
  if (bi == -1 && atomicCAS(&W->bigIdx, -1, -2) == -1)
  {
    int v = atomicAdd(bigCount, 1);
    if (v >= nbig)
    {
      atomicMax(err, ERR_SCRATCH);
      v = nbig - 1;
    }
    __threadfence();
    atomicExch(&W->bigIdx, v);
    return v;
  }
  do
  {
    bi = atomicAdd(&W->bigIdx, 0); // another warp is between its CAS and its exchange: a few cycles
  } while (bi < 0);
  return bi;
}

// Enumerates the set bits of a warp's 32 words (lane i holds word i) so that every lane gets its own
// pixel: returns, for pixel number p (0-based, < total), the lane holding its word and the bit position.
__device__ __forceinline__ void
warp_pixel(uint32_t myWord, int incl, int p, int & srcLane, int & bit)
{
  int lo = 0; // smallest lane whose inclusive count exceeds p
#pragma unroll
  for (int step = 16; step > 0; step >>= 1)
  {
    int probe = lo + step - 1;
    int v = __shfl_sync(0xFFFFFFFFu, incl, probe);
    if (v <= p)
      lo += step;
  }
  lo = min(lo, 31);
  srcLane = lo;
  const uint32_t w = __shfl_sync(0xFFFFFFFFu, myWord, lo);
  const int inclSrc = __shfl_sync(0xFFFFFFFFu, incl, lo);
  const int before = inclSrc - __popc(w);
  bit = __fns(w, 0, max(1, p - before + 1));
}

template <typename T>
__global__ void __launch_bounds__(256)
k_dt_hist(NodeWork * work, const DtItem * __restrict__ items, const int * __restrict__ itemCount, int itemCap,
          const uint32_t * __restrict__ scratch, unsigned * hist, int histStride, int histSide, unsigned * bigHist, int * bigCount,
          int nbig, int * err)
{
  pdl_wait();
  const int lane = threadIdx.x & 31;
  // consecutive items belong to one node and cost about the same (far pixels search many rows): deal them to
  // consecutive CTAs, not to the warps of one CTA, so that no SM collects a run of expensive ones
  const int gw = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, nw = (gridDim.x * blockDim.x) >> 5;
  const int nItems = min(*itemCount, itemCap);
  for (int it = gw; it < nItems; it += nw) // one warp per item
  {
    const DtItem item = items[it];
    NodeWork * W = work + item.node;
    const int pitch = W->pitch, words = W->words, xr0 = W->xr0, xr1 = W->xr1, nodeH = W->h;
    const uint32_t * Xb = scratch + W->scratchOff;
    const uint32_t * Ib = Xb + words;
    const uint32_t * Sb = Ib + words;
    const uint32_t * rowInfo = Sb + words;
    unsigned * h0 = hist + (size_t)item.node * histStride;
    const int wBeg = item.w0;
    // lane i < DT_HCHUNK owns word wBeg+i; the pixels of the chunk are dealt out one per lane
    const int myIdx = wBeg + lane;
    const bool mine = lane < DT_HCHUNK && myIdx < words;
    const uint32_t myS = mine ? __ldg(Sb + myIdx) : 0u;
    const uint32_t myI = mine ? __ldg(Ib + myIdx) : 0u;
    int incl = __popc(myS);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= o)
        incl += t;
    }
    const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    for (int base = 0; base < total; base += 32)
    {
      const int p = base + lane;
      const bool on = p < total;
      int src, bit;
      warp_pixel(myS, incl, on ? p : total - 1, src, bit);
      const uint32_t iw = __shfl_sync(0xFFFFFFFFu, myI, src);
      int d2 = 0, side = 0;
      bool over = false;
      if (on)
      {
        const int idx = wBeg + src;
        const int r = idx / pitch, wi = idx - r * pitch;
        d2 = dt_edt_sq(Xb, rowInfo, pitch, nodeH, xr0, xr1, wi * 32 + bit, r, 0x7FFFFFFF);
        side = (iw >> bit) & 1u ? 0 : 1;
        // the node's own table is indexed by bin (10 bins per pixel of distance: uchar all 256 wrapped bins,
        // the 16-bit types distances below histSide / 10 pixels); farther pixels go to a wrapped-bin table
        const int direct = dt_bin<T>(d2, err);
        if (direct < histSide)
        {
          atomicAdd(&h0[side * histSide + direct], 1u);
          unsigned * tw = h0 + 2 * histSide + (direct >> 5); // bitmap of touched entries
          if (!(__ldcg(tw) & (1u << (direct & 31))))
            atomicOr(tw, 1u << (direct & 31));
        }
        else
          over = true;
      }
      const unsigned need = __ballot_sync(0xFFFFFFFFu, over);
      if (need)
      {
        const int leader = __ffs(need) - 1;
        int bi = 0;
        if (lane == leader)
          bi = dt_get_big(W, bigCount, nbig, err);
        bi = __shfl_sync(0xFFFFFFFFu, bi, leader);
        if (over)
        {
          const int bin = dt_bin<T>(d2, err);
          unsigned * bh = bigHist + (size_t)bi * DT_BIG_STRIDE;
          atomicAdd(&bh[side * 65536 + bin], 1u);
          unsigned * tw = bh + 2 * 65536 + (bin >> 5);
          if (!(__ldcg(tw) & (1u << (bin & 31))))
            atomicOr(tw, 1u << (bin & 31));
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// One CTA per node.  Both histogram tables carry a bitmap of touched entries; only those can change the
// score (X:481-494), so each thread walks the set bits of its own bitmap words and a block-wide scan
// supplies the running sums: constant depth whatever the table length.
template <typename T>
__global__ void __launch_bounds__(256)
k_dt_scan(NodeWork * work, const int * __restrict__ count, unsigned * hist, int histStride, int histSide, unsigned * bigHist)
{
  pdl_wait();
  __shared__ long long redA[8], redB[8], bestD[8];
  __shared__ int bestI[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = *count;
  for (int nidx = blockIdx.x; nidx < n; nidx += gridDim.x)
  {
    NodeWork * W = work + nidx;
    if (!W->alive)
      continue;
    unsigned * hs = hist + (size_t)nidx * histStride;
    unsigned * hsBm = hs + 2 * histSide;
    const int smallWords = histSide / 32;
    const int bi = W->bigIdx;
    const bool big = bi >= 0;
    unsigned * cntI = hs;
    unsigned * cntJ = hs + histSide;
    unsigned * bm = hsBm;
    int wpt = 1; // bitmap words per thread (contiguous, so thread order = bin order)
    __syncthreads();
    if (big)
    {
      // Some bin exceeded the node's own table: work in the full wrapped-bin table of X:431 and fold the
      // node's table in (same bin numbers).
      unsigned * bh = bigHist + (size_t)bi * DT_BIG_STRIDE;
      unsigned * bbm = bh + 2 * 65536;
      for (int wI = tid; wI < smallWords; wI += 256)
      {
        unsigned bits = hsBm[wI];
        if (!bits)
          continue;
        hsBm[wI] = 0u;
        while (bits)
        {
          int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int k = wI * 32 + b;
          const int bin = k;
          atomicAdd(&bh[bin], hs[k]);
          atomicAdd(&bh[65536 + bin], hs[histSide + k]);
          atomicOr(&bbm[bin >> 5], 1u << (bin & 31));
          hs[k] = 0u;
          hs[histSide + k] = 0u;
        }
      }
      __syncthreads();
      cntI = bh;
      cntJ = bh + 65536;
      bm = bbm;
      wpt = 8;
    }
    const int nWords = big ? 2048 : smallWords;
    unsigned myBits[8];
    long long li = 0, lj = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      const int wI = tid * wpt + q;
      myBits[q] = (q < wpt && wI < nWords) ? bm[wI] : 0u;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
    {
      unsigned bits = myBits[q];
      while (bits)
      {
        int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int k = (tid * wpt + q) * 32 + b;
        li += cntI[k];
        lj += cntJ[k];
      }
    }
    // block exclusive scan of (li, lj)
    long long sa = li, sb = lj;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      long long ta = __shfl_up_sync(0xFFFFFFFFu, sa, o), tb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
      if (lane >= o)
      {
        sa += ta;
        sb += tb;
      }
    }
    if (lane == 31)
    {
      redA[warp] = sa;
      redB[warp] = sb;
    }
    __syncthreads();
    long long oa = 0, ob = 0, iTot = 0, jTot = 0;
    for (int k = 0; k < 8; ++k)
    {
      if (k < warp)
      {
        oa += redA[k];
        ob += redB[k];
      }
      iTot += redA[k];
      jTot += redB[k];
    }
    int thr = -1; // -1: no pixel outside the intersection, the median is the intersection (X:463-466)
    if (iTot + jTot > 0) // block-uniform
    {
      long long si = oa + sa - li, sj = ob + sb - lj;
      long long bd = 0x7FFFFFFFFFFFFFFFll;
      int bidx = 0x7FFFFFFF;
      if (tid == 0 && !(myBits[0] & 1u))
      {
        // bin 0 is always visited by the reference (bestDiff starts at LLONG_MAX, X:482-494)
        bd = llabs(llabs(iTot) - llabs(jTot));
        bidx = 0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
      {
        unsigned bits = myBits[q];
        while (bits)
        {
          int b = __ffs(bits) - 1;
          bits &= bits - 1;
          const int k = (tid * wpt + q) * 32 + b;
          si += cntI[k];
          sj += cntJ[k];
          long long d = llabs(llabs(iTot - si + sj) - llabs(jTot - sj + si));
          if (d < bd)
          {
            bd = d;
            bidx = k;
          }
          cntI[k] = 0u; // leave the tables clean for the next user
          cntJ[k] = 0u;
        }
        if (myBits[q])
          bm[tid * wpt + q] = 0u;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        long long od = __shfl_xor_sync(0xFFFFFFFFu, bd, o);
        int oi = __shfl_xor_sync(0xFFFFFFFFu, bidx, o);
        if (od < bd || (od == bd && oi < bidx))
        {
          bd = od;
          bidx = oi;
        }
      }
      if (lane == 0)
      {
        bestD[warp] = bd;
        bestI[warp] = bidx;
      }
      __syncthreads();
      bd = bestD[0];
      bidx = bestI[0];
      for (int k = 1; k < 8; ++k)
        if (bestD[k] < bd || (bestD[k] == bd && bestI[k] < bidx))
        {
          bd = bestD[k];
          bidx = bestI[k];
        }
      thr = dt_threshold_d2(bidx); // bidx = bestBin (X:481-494)
    }
    if (tid == 0)
      W->thr = thr;
  }
}

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t
bit_range(int lo, int hi) // bits lo..hi (0 <= lo <= hi <= 31)
{
  return (0xFFFFFFFFu << lo) & (0xFFFFFFFFu >> (31 - hi));
}

// With the threshold known, {d2 <= thr} is a union of horizontally dilated intersection rows:
// row r+k contributes its pixels widened by floor(sqrt(thr - k*k)).  Bit-parallel per 32-pixel word.
__global__ void __launch_bounds__(256)
k_dt_median(NodeWork * work, const DtItem * __restrict__ items, const int * __restrict__ itemCount, uint32_t * scratch, int SU0,
            int SU1, int SU2, int * err)
{
  pdl_wait();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nItems = *itemCount;
  for (int it = blockIdx.x; it < nItems; it += gridDim.x)
  {
    const DtItem item = items[it];
    NodeWork * W = work + item.node;
    const int pitch = W->pitch, words = W->words, xr0 = W->xr0, xr1 = W->xr1, thr = W->thr;
    const uint32_t * Xb = scratch + W->scratchOff;
    uint32_t * Sb = scratch + W->scratchOff + 2 * (size_t)words;
    const uint32_t * rowInfo = Sb + words;
    const int axis = W->nd.axis;
    const int SU = axis == 0 ? SU1 : SU0, SV = axis == 2 ? SU1 : SU2; // slice extents (X:681-693)
    const int ux0 = W->ux0, uy0 = W->uy0;
    const int myIdx = item.w0 + warp * 32 + lane;
    int bx0 = 0x7FFFFFFF, bx1 = -0x7FFFFFFF, by0 = 0x7FFFFFFF, by1 = -0x7FFFFFFF;
    if (myIdx < words)
    {
      const uint32_t myS = Sb[myIdx];
      uint32_t m = __ldg(Xb + myIdx);
      const int r = myIdx / pitch, wi = myIdx - r * pitch;
      if (thr >= 0 && myS)
      {
        const int x0 = wi * 32;
        uint32_t pass = 0;
        int R = (int)sqrtf((float)thr) + 1; // widened below to floor(sqrt(thr - k*k)) exactly
        for (int k = 0; k * k <= thr; ++k)
        {
          while (R * R + k * k > thr)
            --R;
#pragma unroll
          for (int sgn = 0; sgn < 2; ++sgn)
          {
            if (sgn && k == 0)
              continue;
            const int rr = sgn ? r + k : r - k;
            if (rr < xr0 || rr > xr1)
              continue;
            const uint32_t info = __ldg(rowInfo + rr);
            if (!info)
              continue;
            const int a = info & 0x1FFF, b = (info >> 13) & 0x1FFF;
            if (a - R > x0 + 31 || b + R < x0)
              continue;
            if ((info >> 26) & 1u)
              pass |= bit_range(max(a - R, x0) - x0, min(b + R, x0 + 31) - x0);
            else
            {
              // several runs in this row: widen each piece that can reach the word
              const uint32_t * row = Xb + (size_t)rr * pitch;
              const int wlo = max(0, (x0 - R) >> 5), whi = min(pitch - 1, (x0 + 31 + R) >> 5);
              for (int ws = wlo; ws <= whi; ++ws)
              {
                uint32_t wv = __ldg(row + ws);
                while (wv)
                {
                  const int sbit = __ffs(wv) - 1;
                  const uint32_t t = ~(wv >> sbit);
                  const int len = t ? __ffs(t) - 1 : 32 - sbit;
                  const int lo = ws * 32 + sbit - R, hi = ws * 32 + sbit + len - 1 + R;
                  if (hi >= x0 && lo <= x0 + 31)
                    pass |= bit_range(max(lo, x0) - x0, min(hi, x0 + 31) - x0);
                  wv = (len >= 32) ? 0u : (wv & ~(((1u << len) - 1u) << sbit));
                }
              }
            }
          }
          if ((pass & myS) == myS)
            break;
        }
        m |= pass & myS;
      }
      Sb[myIdx] = m;
      // bounding box of the part that lies inside the slice (X:679-712)
      const int y = uy0 + r;
      if (m && y >= 0 && y < SV)
      {
        const int xa = ux0 + 32 * wi;
        uint32_t c = m;
        if (xa < 0)
          c = (xa <= -32) ? 0u : (c & (0xFFFFFFFFu << (-xa)));
        if (xa + 32 > SU)
          c = (xa >= SU) ? 0u : (c & (0xFFFFFFFFu >> (xa + 32 - SU)));
        if (c)
        {
          bx0 = xa + __ffs(c) - 1;
          bx1 = xa + 31 - __clz(c);
          by0 = by1 = y;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      bx0 = min(bx0, __shfl_xor_sync(0xFFFFFFFFu, bx0, o));
      by0 = min(by0, __shfl_xor_sync(0xFFFFFFFFu, by0, o));
      bx1 = max(bx1, __shfl_xor_sync(0xFFFFFFFFu, bx1, o));
      by1 = max(by1, __shfl_xor_sync(0xFFFFFFFFu, by1, o));
    }
    if (lane == 0 && by1 >= by0)
    {
      atomicMin(&W->bx0, bx0);
      atomicMax(&W->bx1, bx1);
      atomicMin(&W->by0, by0);
      atomicMax(&W->by1, by1);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
__global__ void
k_dt_alloc(NodeWork * work, const int * __restrict__ count, unsigned long long * arenaTop, unsigned long long arenaCap, Node * next,
           int * nextCount, int nextCap, DtItem * witems, int * witemCount, int witemCap, EmitRec * emit, int * emitCount,
           int emitCap, unsigned long long emitBase, int * err)
{
  pdl_wait();
  const int nidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (nidx >= *count)
    return;
  NodeWork * W = work + nidx;
  if (!W->alive)
    return;
  if (W->by1 < W->by0)
  {
    W->alive = 0; // nothing of the median lies inside the image
    return;
  }
  MaskRef M;
  M.x0 = W->bx0;
  M.y0 = W->by0;
  M.w = W->bx1 - W->bx0 + 1;
  M.h = W->by1 - W->by0 + 1;
  M.pitch = (M.w + 31) >> 5;
  M.pad = 0;
  const int mwords = M.h * M.pitch;
  M.off = atomicAdd(arenaTop, (unsigned long long)mwords);
  if (M.off + (unsigned long long)mwords > arenaCap)
  {
    atomicMax(err, ERR_ARENA);
    W->alive = 0;
    return;
  }
  W->M = M;
  if (emit)
  {
    const int e = atomicAdd(emitCount, 1);
    if (e < emitCap)
    {
      EmitRec r;
      r.M = M;
      r.M.off = M.off - emitBase;
      r.axis = W->nd.axis;
      r.label = W->nd.label;
      r.mid = W->mid;
      r.pad = 0;
      emit[e] = r;
      if (W->nd.slot0 >= 0)
        emit_slot_table(emit, emitCap)[W->nd.slot0 + W->mid - W->nd.lo - 1] = e;
    }
    else
      atomicMax(err, ERR_NODE_LIST);
  }
  const int nItems = (mwords + DT_CHUNK - 1) / DT_CHUNK;
  const int base = atomicAdd(witemCount, nItems);
  if (base + nItems > witemCap)
  {
    atomicMax(err, ERR_SCRATCH);
    W->alive = 0;
    return;
  }
  for (int q = 0; q < nItems; ++q)
  {
    DtItem it;
    it.node = nidx;
    it.w0 = q * DT_CHUNK;
    witems[base + q] = it;
  }
  // children (X:766-792)
  const Node & nd = W->nd;
  const int mid = W->mid;
  if (abs(nd.i - nd.j) > 2)
  {
    const bool first = abs(nd.i - mid) > 1, second = abs(nd.j - mid) > 1;
    const int nnew = (first ? 1 : 0) + (second ? 1 : 0);
    if (nnew)
    {
      int pos = atomicAdd(nextCount, nnew);
      if (pos + nnew > nextCap)
      {
        atomicMax(err, ERR_NODE_LIST);
        return;
      }
      if (first)
      {
        Node c;
        c.A = nd.A;
        c.B = M;
        c.tx = W->iTx;
        c.ty = W->iTy;
        c.i = nd.i;
        c.j = mid;
        c.axis = nd.axis;
        c.label = nd.label;
        c.slot0 = nd.slot0;
        c.lo = nd.lo;
        next[pos++] = c;
      }
      if (second)
      {
        Node c;
        c.A = nd.B;
        c.B = M;
        c.tx = W->jTx;
        c.ty = W->jTy;
        c.i = nd.j;
        c.j = mid;
        c.axis = nd.axis;
        c.label = nd.label;
        c.slot0 = nd.slot0;
        c.lo = nd.lo;
        next[pos++] = c;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// max rule on VPQ voxels at once (one aligned 8-byte group): out = max(out, label) where the bit is set and
// the input voxel is 0 (X:754-763 with the final overwrite of X:1623-1636 folded in)
template <typename T>
__device__ __forceinline__ void
write_group(const T * __restrict__ in, T * out, unsigned long long g /* first voxel, multiple of VPQ */, unsigned bits, T label)
{
  constexpr int VPQ = 8 / sizeof(T);
  union Q
  {
    unsigned long long u;
    T v[VPQ];
  };
  Q qi, qo, qn;
  qi.u = __ldg(reinterpret_cast<const unsigned long long *>(in + g));
  unsigned long long * po = reinterpret_cast<unsigned long long *>(out + g);
  qo.u = *po;
  for (;;)
  {
    qn = qo;
#pragma unroll
    for (int k = 0; k < VPQ; ++k)
      if (((bits >> k) & 1u) && qi.v[k] == 0 && qo.v[k] < label)
        qn.v[k] = label;
    if (qn.u == qo.u)
      return;
    unsigned long long prev = atomicCAS(po, qo.u, qn.u);
    if (prev == qo.u)
      return;
    qo.u = prev;
  }
}

// same with the two groups already loaded (lets a caller keep several groups in flight)
template <typename T>
__device__ __forceinline__ void
write_group_loaded(T * out, unsigned long long g, unsigned bits, T label, unsigned long long qinU, unsigned long long qoutU)
{
  constexpr int VPQ = 8 / sizeof(T);
  union Q
  {
    unsigned long long u;
    T v[VPQ];
  };
  Q qi, qo, qn;
  qi.u = qinU;
  qo.u = qoutU;
  unsigned long long * po = reinterpret_cast<unsigned long long *>(out + g);
  for (;;)
  {
    qn = qo;
#pragma unroll
    for (int k = 0; k < VPQ; ++k)
      if (((bits >> k) & 1u) && qi.v[k] == 0 && qo.v[k] < label)
        qn.v[k] = label;
    if (qn.u == qo.u)
      return;
    unsigned long long prev = atomicCAS(po, qo.u, qn.u);
    if (prev == qo.u)
      return;
    qo.u = prev;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_dt_emit(const NodeWork * __restrict__ work, const DtItem * __restrict__ witems, const int * __restrict__ witemCount,
          const uint32_t * __restrict__ scratch, uint32_t * arena, const T * __restrict__ in, T * out, long long VX, long long VY,
          long long VZ)
{
  pdl_wait();
  constexpr int VPQ = 8 / sizeof(T);       // voxels per aligned 8-byte group
  constexpr int GPW = 32 / VPQ + 1;        // groups a 32-pixel word can touch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nItems = *witemCount;
  const unsigned long long nvox = (unsigned long long)VX * VY * VZ;
  for (int it = blockIdx.x; it < nItems; it += gridDim.x)
  {
    const DtItem item = witems[it];
    const NodeWork * W = work + item.node;
    const MaskRef M = W->M;
    const int pitch = W->pitch, ux0 = W->ux0, uy0 = W->uy0, mid = W->mid, axis = W->nd.axis;
    const T label = (T)W->nd.label;
    const uint32_t * Sb = scratch + W->scratchOff + 2 * (size_t)W->words;
    long long su, sv;
    unsigned long long vbase;
    if (axis == 2)
    {
      su = 1;
      sv = VX;
      vbase = (unsigned long long)mid * VX * VY;
    }
    else if (axis == 1)
    {
      su = 1;
      sv = VX * VY;
      vbase = (unsigned long long)mid * VX;
    }
    else
    {
      su = VX;
      sv = VX * VY;
      vbase = (unsigned long long)mid;
    }
    const int mwords = M.h * M.pitch;
    const int wBeg = item.w0 + warp * 32, wEnd = min(mwords, wBeg + 32);
    if (wBeg >= mwords)
      continue;
    // lane i crops word wBeg+i of the mid shape out of the median rows and stores it
    uint32_t myM = 0;
    {
      const int cell = wBeg + lane;
      if (cell < wEnd)
      {
        const int r = cell / M.pitch, wi = cell - r * M.pitch;
        const int y = M.y0 + r, x = M.x0 + 32 * wi;
        const uint32_t * row = Sb + (size_t)(y - uy0) * pitch;
        const int rx = x - ux0; // >= 0
        const int sw = rx >> 5, shf = rx & 31;
        const uint32_t lo = __ldg(row + sw);
        const uint32_t hi = sw + 1 < pitch ? __ldg(row + sw + 1) : 0u;
        myM = __funnelshift_r(lo, hi, shf);
        const int valid = M.w - 32 * wi;
        if (valid < 32)
          myM &= (1u << valid) - 1u;
        arena[M.off + cell] = myM;
      }
    }
    if (su == 1)
    {
      // x runs along the bits: aligned 8-byte groups, one per lane
      const int nTasks = (wEnd - wBeg) * GPW;
      for (int tb = 0; tb < nTasks; tb += 32)
      {
        const int task = tb + lane;
        const int c = min(task / GPW, wEnd - wBeg - 1), q = task % GPW;
        const uint32_t m = __shfl_sync(0xFFFFFFFFu, myM, c);
        if (task < nTasks && m)
        {
          const int cell = wBeg + c;
          const int r = cell / M.pitch, wi = cell - r * M.pitch;
          const unsigned long long a0 = vbase + (unsigned long long)((long long)(M.x0 + 32 * wi) + (long long)(M.y0 + r) * sv);
          const unsigned long long g = (a0 & ~(unsigned long long)(VPQ - 1)) + (unsigned long long)q * VPQ;
          // bits of this group: voxel g+k <-> bit (g + k - a0)
          const long long sh = (long long)g - (long long)a0; // in (-VPQ, 32]
          unsigned bits;
          if (sh >= 32)
            bits = 0u;
          else if (sh >= 0)
            bits = (m >> sh) & ((1u << VPQ) - 1u);
          else
            bits = (m << (-sh)) & ((1u << VPQ) - 1u);
          if (bits && g + VPQ <= nvox)
            write_group<T>(in, out, g, bits, label);
          else if (bits)
          {
            for (int k = 0; k < VPQ; ++k)
              if (((bits >> k) & 1u) && g + k < nvox && in[g + k] == 0)
                atomic_max_pixel(out + g + k, label);
          }
        }
      }
    }
    else if (W->nd.slot0 < 0) // registered axis-0 trees are written x-run by x-run after the last level (k_apply_axis0)
    {
      for (int c = 0; c < wEnd - wBeg; ++c)
      {
        const uint32_t m = __shfl_sync(0xFFFFFFFFu, myM, c);
        if ((m >> lane) & 1u)
        {
          const int cell = wBeg + c;
          const int r = cell / M.pitch, wi = cell - r * M.pitch;
          const unsigned long long a =
            vbase + (unsigned long long)((long long)(M.x0 + 32 * wi + lane) * su + (long long)(M.y0 + r) * sv);
          if (in[a] == 0)
            atomic_max_pixel(out + a, label);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Replays the volume writes of mid slices produced elsewhere (another GPU's shard): same write rule as k_dt_emit.
template <typename T>
__global__ void __launch_bounds__(256)
k_apply_masks(const EmitRec * __restrict__ recs, int nrecsHost, const int * __restrict__ nrecsDev, const uint32_t * __restrict__ words,
              const T * __restrict__ in, T * out, long long VX, long long VY, long long VZ, int skipAxis0)
{
  pdl_wait();
  const int nrecs = nrecsDev ? min(*nrecsDev, nrecsHost) : nrecsHost; // device-side count capped by the list capacity
  constexpr int VPQ = 8 / sizeof(T);
  constexpr int GPW = 32 / VPQ + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned long long nvox = (unsigned long long)VX * VY * VZ;
  for (int ri = blockIdx.x; ri < nrecs; ri += gridDim.x)
  {
    const EmitRec R = recs[ri];
    if (skipAxis0 && R.axis == 0)
      continue; // written by k_apply_axis0
    const MaskRef M = R.M;
    const T label = (T)R.label;
    long long su, sv;
    unsigned long long vbase;
    if (R.axis == 2)
    {
      su = 1;
      sv = VX;
      vbase = (unsigned long long)R.mid * VX * VY;
    }
    else if (R.axis == 1)
    {
      su = 1;
      sv = VX * VY;
      vbase = (unsigned long long)R.mid * VX;
    }
    else
    {
      su = VX;
      sv = VX * VY;
      vbase = (unsigned long long)R.mid;
    }
    if (su == 1)
    {
      // x runs along the bits.  A warp takes one mask row at a time and three of its words per iteration: lane
      // 9*s + q handles aligned 8-byte group q of word s (27 lanes busy), so no per-task index arithmetic is left.
      const int sub = lane / GPW, q = lane - sub * GPW;
      for (int r = warp; r < M.h; r += nwarps)
      {
        const unsigned long long rowBase = vbase + (unsigned long long)((long long)M.x0 + (long long)(M.y0 + r) * sv);
        const uint32_t * rowWords = words + M.off + (size_t)r * M.pitch;
        for (int wi0 = 0; wi0 < M.pitch; wi0 += 3)
        {
          const int wi = wi0 + sub;
          if (sub >= 3 || wi >= M.pitch)
            continue;
          const uint32_t m = __ldg(rowWords + wi);
          if (!m)
            continue;
          const unsigned long long a0 = rowBase + 32ull * wi;
          const unsigned long long g = (a0 & ~(unsigned long long)(VPQ - 1)) + (unsigned long long)q * VPQ;
          const int sh = (int)((long long)g - (long long)a0); // in (-VPQ, 32]
          unsigned bits;
          if (sh >= 32)
            bits = 0u;
          else if (sh >= 0)
            bits = (m >> sh) & ((1u << VPQ) - 1u);
          else
            bits = (m << (-sh)) & ((1u << VPQ) - 1u);
          if (!bits)
            continue;
          if (g + VPQ <= nvox)
            write_group<T>(in, out, g, bits, label);
          else
            for (int k = 0; k < VPQ; ++k)
              if (((bits >> k) & 1u) && g + k < nvox && in[g + k] == 0)
                atomic_max_pixel(out + g + k, label);
        }
      }
    }
    else
    {
      const int mwords = M.h * M.pitch;
      for (int cb = warp * 32; cb < mwords; cb += nwarps * 32)
      {
        const int cEnd = min(mwords, cb + 32);
        const uint32_t myM = cb + lane < cEnd ? __ldg(words + M.off + cb + lane) : 0u;
        for (int c = 0; c < cEnd - cb; ++c)
        {
          const uint32_t m = __shfl_sync(0xFFFFFFFFu, myM, c);
          if ((m >> lane) & 1u)
          {
            const int cell = cb + c;
            const int r = cell / M.pitch, wi = cell - r * M.pitch;
            const unsigned long long a =
              vbase + (unsigned long long)((long long)(M.x0 + 32 * wi + lane) * su + (long long)(M.y0 + r) * sv);
            if (in[a] == 0)
              atomic_max_pixel(out + a, label);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Volume writes of the trees whose slices are perpendicular to x (axis 0).  Written slice by slice, every pixel of
// such a mid slice is a 2-byte read-modify-write X voxels away from the next: one 32-byte sector per voxel.  But the
// mid slices of ONE tree sit at consecutive x, so for a fixed (y, z) their pixels form a contiguous x-run.  One CTA
// takes up to 32 consecutive slice positions of a tree (Axis0Group; the slot table maps position -> record); a warp
// takes a 32-pixel cell of the union region, lane k fetches that cell's word of slice k (funnel-shifted to the
// cell), the 32 x n bit matrix is transposed with shuffles, and lane b writes the x-run of pixel b in aligned
// 8-byte groups with the max rule (X:754-763).
template <typename T>
__global__ void __launch_bounds__(256)
k_apply_axis0(const Axis0Group * __restrict__ groups, int ngroups, int parts, const EmitRec * __restrict__ recs,
              const int * __restrict__ slots, const uint32_t * __restrict__ words, const T * __restrict__ in, T * out, long long VX,
              long long VY, long long VZ)
{
  pdl_wait();
  constexpr int VPQ = 8 / sizeof(T);
  __shared__ EmitRec s_rec[32];
  __shared__ int s_box[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const unsigned long long nvox = (unsigned long long)VX * VY * VZ;
  // work unit = (group, part): the cells of a group are dealt to `parts` CTAs, there are far fewer trees than SMs x 8
  for (int unit = blockIdx.x; unit < ngroups * parts; unit += gridDim.x)
  {
    const int gi = unit / parts, part = unit - gi * parts;
    const Axis0Group G = groups[gi];
    __syncthreads(); // previous group done with s_rec / s_box
    if (threadIdx.x < 4)
      s_box[threadIdx.x] = threadIdx.x < 2 ? 0x7FFFFFFF : -0x7FFFFFFF;
    __syncthreads();
    if (threadIdx.x < 32)
    {
      EmitRec R{};
      const int e = threadIdx.x < G.n ? slots[G.slot0 + threadIdx.x] : -1;
      if (e >= 0)
        R = recs[e];
      else
        R.axis = -1;
      s_rec[threadIdx.x] = R;
      if (e >= 0 && R.M.w > 0 && R.M.h > 0)
      {
        atomicMin(&s_box[0], R.M.x0);
        atomicMin(&s_box[1], R.M.y0);
        atomicMax(&s_box[2], R.M.x0 + R.M.w);
        atomicMax(&s_box[3], R.M.y0 + R.M.h);
      }
    }
    __syncthreads();
    const int U0 = s_box[0], V0 = s_box[1], U1 = s_box[2], V1 = s_box[3];
    if (U1 <= U0 || V1 <= V0)
      continue;
    // my slice (lane k): geometry in registers
    const EmitRec mine = s_rec[lane];
    const bool valid = mine.axis == 0 && mine.M.w > 0;
    // x of slot 0 and the label, from any valid record
    const unsigned vmask = __ballot_sync(0xFFFFFFFFu, valid);
    const int src = __ffs(vmask) - 1;
    const int xSlot0 = __shfl_sync(0xFFFFFFFFu, mine.mid - lane, src);
    const T label = (T)__shfl_sync(0xFFFFFFFFu, mine.label, src);
    const uint32_t * myWords = words + mine.M.off;
    const int ncw = (U1 - U0 + 31) >> 5, ncells = ncw * (V1 - V0);
    for (int cell = part * nwarps + warp; cell < ncells; cell += parts * nwarps)
    {
      const int v = V0 + cell / ncw, Uc = U0 + 32 * (cell % ncw);
      uint32_t W = 0u;
      if (valid && v >= mine.M.y0 && v < mine.M.y0 + mine.M.h)
      {
        const uint32_t * row = myWords + (size_t)(v - mine.M.y0) * mine.M.pitch;
        const int rx = Uc - mine.M.x0; // column of the mask that bit 0 of the cell maps to
        if (rx >= 0)
        {
          const int sw = rx >> 5, sh = rx & 31;
          const uint32_t lo = sw < mine.M.pitch ? __ldg(row + sw) : 0u;
          const uint32_t hi = sw + 1 < mine.M.pitch ? __ldg(row + sw + 1) : 0u;
          W = __funnelshift_r(lo, hi, sh);
        }
        else if (rx > -32)
          W = __ldg(row) << (-rx);
      }
      unsigned any = __ballot_sync(0xFFFFFFFFu, W != 0u);
      if (!any)
        continue;
      // transpose: bit k of `run` = pixel `lane` of slice k
      unsigned run = 0u;
      while (any)
      {
        const int k = __ffs(any) - 1;
        any &= any - 1;
        const uint32_t Wk = __shfl_sync(0xFFFFFFFFu, W, k);
        run |= ((Wk >> lane) & 1u) << k;
      }
      if (!run)
        continue;
      // pixel (u = y, v = z): voxels x = xSlot0 + k for the set bits k
      const unsigned long long a0 = ((unsigned long long)v * VY + (unsigned long long)(Uc + lane)) * VX + (unsigned long long)xSlot0;
      const int kLast = 31 - __clz(run);
      for (unsigned long long g = (a0 + (__ffs(run) - 1)) & ~(unsigned long long)(VPQ - 1); g <= a0 + kLast; g += VPQ)
      {
        const long long sh = (long long)g - (long long)a0; // in (-VPQ, 32)
        const unsigned bits = (sh >= 0 ? (run >> sh) : (run << (-sh))) & ((1u << VPQ) - 1u);
        if (!bits)
          continue;
        if (g + VPQ <= nvox)
          write_group<T>(in, out, g, bits, label);
        else
          for (int k = 0; k < VPQ; ++k)
            if (((bits >> k) & 1u) && g + k < nvox && in[g + k] == 0)
              atomic_max_pixel(out + g + k, label);
      }
    }
  }
}

void
launch_apply_masks(Launch & L, int ptype, const EmitRec * recs, int nrecs, const int * nrecsDev, const uint32_t * words, const void * in,
                   void * out, const long long dims[3], int skipAxis0)
{
  if (nrecs <= 0)
    return;
  const int grid = std::max(1, std::min(nrecs, L.sms * 8));
  L.pre("k_apply_masks");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_apply_masks<uint8_t>, grid, 256, 0, L.stream, recs, nrecs, nrecsDev, words, (const uint8_t *)in, (uint8_t *)out, dims[0], dims[1],
                                                         dims[2], skipAxis0);
      break;
    case MCI_U16:
      launch_pdl(k_apply_masks<uint16_t>, grid, 256, 0, L.stream, recs, nrecs, nrecsDev, words, (const uint16_t *)in, (uint16_t *)out, dims[0],
                                                           dims[1], dims[2], skipAxis0);
      break;
    default:
      launch_pdl(k_apply_masks<int16_t>, grid, 256, 0, L.stream, recs, nrecs, nrecsDev, words, (const int16_t *)in, (int16_t *)out, dims[0], dims[1],
                                                          dims[2], skipAxis0);
  }
  L.post("k_apply_masks");
}

void
launch_apply_axis0(Launch & L, int ptype, const Axis0Group * groups, int ngroups, const EmitRec * recs, const int * slots,
                   const uint32_t * words, const void * in, void * out, const long long dims[3])
{
  if (ngroups <= 0)
    return;
  const int parts = std::max(1, std::min(16, (L.sms * 8 + ngroups - 1) / ngroups));
  const int grid = std::max(1, std::min(ngroups * parts, L.sms * 8));
  L.pre("k_apply_axis0");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_apply_axis0<uint8_t>, grid, 256, 0, L.stream, groups, ngroups, parts, recs, slots, words, (const uint8_t *)in, (uint8_t *)out, dims[0],
                                                         dims[1], dims[2]);
      break;
    case MCI_U16:
      launch_pdl(k_apply_axis0<uint16_t>, grid, 256, 0, L.stream, groups, ngroups, parts, recs, slots, words, (const uint16_t *)in, (uint16_t *)out, dims[0],
                                                          dims[1], dims[2]);
      break;
    default:
      launch_pdl(k_apply_axis0<int16_t>, grid, 256, 0, L.stream, groups, ngroups, parts, recs, slots, words, (const int16_t *)in, (int16_t *)out, dims[0],
                                                         dims[1], dims[2]);
  }
  L.post("k_apply_axis0");
}

// ---------------------------------------------------------------------------------------------------
// ===================================================================================================
// k_dt_tree: every Interpolate1to1 call of every recursion tree (X:556-793, FindMedianImageDistances X:405-516) in ONE
// persistent launch instead of six grid-wide launches per level.  A CTA processes one node from start to end -- prep,
// d2 histograms, scan, median, mid shape, volume write -- and then publishes the node's children (X:766-792) on a
// global ready queue, from which idle CTAs take them: no level barrier anywhere, a tree is as deep as its longest
// path and the GPU stays full whatever the mix of gaps.  The node's bit rows (intersection | i | i xor j -> median)
// and row extents live in shared memory when they fit (else in this CTA's slice of global scratch), the two
// histograms always do.  Mid shapes go to the arena as before; a child may run on another SM, so the parent fences
// after storing its shape and before publishing, and shapes are read with L2-only loads (a 128-byte L1 line can hold
// words of a neighbouring shape that was not written yet when the line was first fetched).  Everything downstream
// (emit records, axis-0 slot table, multi-GPU replay) is unchanged.
// Queue: slots [0, nRoots) are the root nodes (ready from the start); children go to slot nRoots + atomicAdd(tail).
// A CTA claims slot atomicAdd(head) and, if the slot has not been produced yet, waits for its flag -- or for the end:
// `pushed` counts published children, `done` finished nodes; all work is over when done == nRoots + pushed (read
// `done` first: a node publishes its children before it counts as done).
// ===================================================================================================
// 32 bits of row y of a shape starting at slice column X, read from L2 (see k_dt_tree: shapes written by other SMs)
__device__ __forceinline__ uint32_t
mask_word_cg(const uint32_t * data, const MaskRef & m, int y, int X)
{
  int ry = y - m.y0;
  if ((unsigned)ry >= (unsigned)m.h)
    return 0u;
  int rx = X - m.x0;
  int wi = rx >> 5;
  int sh = rx & 31;
  const uint32_t * row = data + (size_t)ry * m.pitch;
  uint32_t lo = ((unsigned)wi < (unsigned)m.pitch) ? __ldcg(row + wi) : 0u;
  uint32_t hi = ((unsigned)(wi + 1) < (unsigned)m.pitch) ? __ldcg(row + wi + 1) : 0u;
  return __funnelshift_r(lo, hi, sh);
}

#define TREE_NODE_WORDS 12288 // shared-memory budget (words) for one node's rows: 3 * words + h + (h + 7) / 8
#define TREE_GROUPS 512       // nodes of up to 16 384 words have their xor pixels dealt across the whole CTA
#define TREE_HWORDS 8         // words per warp visit in the histogram phase (pixels are dealt one per lane); measured
                              // k_dt_tree C3 / C5 in ms: 2 words 0.43 / 6.8, 4: 0.35 / 5.7, 8: 0.35 / 4.9, 32: 0.62 / 4.7

__device__ __forceinline__ int
tree_get_big(int * sBig, int * bigCount, int nbig, int * err)
{
  int bi = atomicAdd(sBig, 0);
  if (bi >= 0)
    return bi;
  if (bi == -1 && atomicCAS(sBig, -1, -2) == -1)
  {
    int v = atomicAdd(bigCount, 1);
    if (v >= nbig)
    {
      atomicMax(err, ERR_SCRATCH);
      v = nbig - 1;
    }
    __threadfence_block();
    atomicExch(sBig, v);
    return v;
  }
  do
  {
    bi = atomicAdd(sBig, 0); // another warp of this CTA is between its CAS and its exchange
  } while (bi < 0);
  return bi;
}


// TREE_THREADS = 512: two CTAs per SM (throughput: many more nodes than CTAs); 1024: one CTA per SM, every phase of a
// node twice as wide (latency: few nodes, the launch lasts as long as the deepest chain of its biggest nodes)
template <typename T, int TREE_THREADS>
__global__ void __launch_bounds__(TREE_THREADS, TREE_THREADS == 512 ? 2 : 1)
k_dt_tree(const TreeParams P)
{
  constexpr int TREE_WARPS = TREE_THREADS / 32;
  pdl_wait();
  extern __shared__ uint32_t tsm[];
  __shared__ __align__(16) Node s_child[2];
  __shared__ __align__(16) Node s_nd;
  __shared__ MaskRef s_M;
  __shared__ int s_nchild, s_slot, s_r0, s_r1, s_big, s_thr, s_alive, s_bx0, s_bx1, s_by0, s_by1, s_maxBin;
  __shared__ int s_gcum[TREE_GROUPS + 1]; // histogram phase: xor pixels before each 32-word group of the node's rows
  __shared__ long long redA[TREE_WARPS], redB[TREE_WARPS], bestD[TREE_WARPS];
  __shared__ int bestI[TREE_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int histSide = P.histSide, smallWords = histSide / 32;
  unsigned * hs = tsm;                       // [2 * histSide] counts of both sides
  unsigned * hsBm = hs + 2 * histSide;       // [histSide / 32] bitmap of touched bins
  uint32_t * nodeSm = hsBm + smallWords;     // [TREE_NODE_WORDS]
  uint32_t * nodeGl = P.scratch + (size_t)blockIdx.x * P.scratchPerCta;
  const T * in = (const T *)P.in;
  T * out = (T *)P.out;
  const int nRoots = *P.rootCount;
#ifdef MCI_TREE_PROFILE
  long long phT[7] = { 0, 0, 0, 0, 0, 0, 0 }, phLast = clock64();
  int phCur = 6;
#define PH(k)                                                                                                                   \
  do                                                                                                                            \
  {                                                                                                                             \
    const long long now_ = clock64();                                                                                           \
    phT[phCur] += now_ - phLast;                                                                                                \
    phLast = now_;                                                                                                              \
    phCur = k;                                                                                                                  \
  } while (0)
#else
#define PH(k) (void)0
#endif
  for (int k = tid; k < 2 * histSide + smallWords; k += TREE_THREADS)
    hs[k] = 0u;
  if (tid == 0)
    s_big = -1;
  __syncthreads();
  volatile int * vflags = P.flags;
  volatile int * vctr = P.rootNext;
  for (;;)
  {
    {
      __syncthreads();
      if (tid == 0)
      {
        // claim a slot; wait until it has been produced, or until no work is left anywhere
        const int slot = atomicAdd(P.rootNext, 1);
        int got = slot < nRoots ? 1 : 0;
        while (!got)
        {
          if (slot - nRoots < P.queueCap && vflags[slot - nRoots])
          {
            got = 1;
            break;
          }
          const int d = vctr[4];
          __threadfence();
          const int pu = vctr[3];
          if (d == nRoots + pu)
            break;
          __nanosleep(200);
        }
        if (got)
        {
          __threadfence();
          const int4 * src = reinterpret_cast<const int4 *>(P.roots + slot);
          int4 * dst = reinterpret_cast<int4 *>(&s_nd);
#pragma unroll
          for (int k = 0; k < (int)(sizeof(Node) / 16); ++k)
            dst[k] = __ldcg(src + k);
        }
        s_slot = got ? slot : -1;
        s_nchild = 0;
        s_r0 = 0x7FFFFFFF;
        s_r1 = -1;
        s_bx0 = s_by0 = 0x7FFFFFFF;
        s_bx1 = s_by1 = -0x7FFFFFFF;
      }
      __syncthreads();
      if (s_slot < 0)
        break;
    }
    do
    {
      const Node nd = s_nd;
      PH(0);
      // ---- prep: translation halving with carry, union region, translate + binarise + AND (X:570-666)
      int t[2] = { nd.tx, nd.ty }, iT[2], jT[2];
      bool carry = false;
#pragma unroll
      for (int d = 0; d < 2; ++d)
      {
        if (!carry)
        {
          iT[d] = t[d] / 2;
          carry = (t[d] % 2) != 0;
        }
        else if (t[d] % 2 == 0)
          iT[d] = t[d] / 2;
        else
        {
          iT[d] = t[d] > 0 ? t[d] / 2 + 1 : t[d] / 2 - 1;
          carry = false;
        }
        jT[d] = iT[d] - t[d];
      }
      const int mid = (nd.i + nd.j + (carry ? 1 : 0)) / 2;
      const int ux0 = min(nd.A.x0 + iT[0], nd.B.x0 + jT[0]), uy0 = min(nd.A.y0 + iT[1], nd.B.y0 + jT[1]);
      const int ux1 = max(nd.A.x0 + nd.A.w + iT[0], nd.B.x0 + nd.B.w + jT[0]);
      const int uy1 = max(nd.A.y0 + nd.A.h + iT[1], nd.B.y0 + nd.B.h + jT[1]);
      const int w = ux1 - ux0, h = uy1 - uy0;
      const int pitch = (w + 31) >> 5;
      const int words = h * pitch;
      const unsigned long long need = 3ull * (unsigned long long)words + (unsigned long long)h + (unsigned long long)((h + 7) / 8);
      if (need > P.scratchPerCta || w > 8191)
      {
        if (tid == 0)
          atomicMax(P.err, ERR_SCRATCH);
        break;
      }
      uint32_t * Xb = need <= TREE_NODE_WORDS ? nodeSm : nodeGl;
      uint32_t * Ib = Xb + words;
      uint32_t * Sb = Ib + words;
      uint32_t * rowInfo = Sb + words;
      const uint32_t * dataA = P.arena + nd.A.off;
      const uint32_t * dataB = P.arena + nd.B.off;
      {
        const int GW = pitch <= 1 ? 1 : (pitch <= 2 ? 2 : (pitch <= 4 ? 4 : (pitch <= 8 ? 8 : (pitch <= 16 ? 16 : 32))));
        const int RPW = 32 / GW, sub = lane / GW, sl = lane - sub * GW;
        int r0 = 0x7FFFFFFF, r1 = -1;
        for (int rb = warp * RPW; rb < h; rb += TREE_WARPS * RPW)
        {
          const int r = rb + sub;
          const bool rowOk = r < h;
          const int y = uy0 + r;
          int first = 0x7FFFFFFF, last = -1, runs = 0;
          uint32_t prevTop = 0;
          for (int wb = 0; wb < pitch; wb += GW)
          {
            const int wi = wb + sl;
            uint32_t a = 0, b = 0;
            if (rowOk && wi < pitch)
            {
              const int x = ux0 + 32 * wi;
              a = mask_word_cg(dataA, nd.A, y - iT[1], x - iT[0]);
              b = mask_word_cg(dataB, nd.B, y - jT[1], x - jT[0]);
              const size_t idx = (size_t)r * pitch + wi;
              Xb[idx] = a & b;
              Ib[idx] = a;
              Sb[idx] = a ^ b;
            }
            const uint32_t xw = a & b;
            uint32_t below = __shfl_up_sync(0xFFFFFFFFu, xw, 1, GW);
            if (sl == 0)
              below = prevTop;
            const uint32_t starts = xw & ~((xw << 1) | (below >> 31));
            runs += __popc(starts);
            if (xw)
            {
              first = min(first, 32 * wi + __ffs(xw) - 1);
              last = max(last, 32 * wi + 31 - __clz(xw));
            }
            prevTop = __shfl_sync(0xFFFFFFFFu, xw, GW - 1, GW);
          }
          for (int o = GW >> 1; o > 0; o >>= 1)
          {
            first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
            last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
            runs += __shfl_xor_sync(0xFFFFFFFFu, runs, o);
          }
          if (sl == 0 && rowOk)
          {
            rowInfo[r] = last >= 0 ? pack_row(first, last, runs == 1) : 0u;
            if (last >= 0)
            {
              r0 = min(r0, r);
              r1 = max(r1, r);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          r0 = min(r0, __shfl_xor_sync(0xFFFFFFFFu, r0, o));
          r1 = max(r1, __shfl_xor_sync(0xFFFFFFFFu, r1, o));
        }
        if (lane == 0 && r1 >= 0)
        {
          atomicMin(&s_r0, r0);
          atomicMax(&s_r1, r1);
        }
      }
      __syncthreads();
      for (int bnd = tid; bnd < (h + 7) / 8; bnd += TREE_THREADS)
      {
        int a = 0x7FFFFFFF, b = -1;
        for (int q = 0; q < 8 && bnd * 8 + q < h; ++q)
        {
          const uint32_t info = rowInfo[bnd * 8 + q];
          if (info)
          {
            a = min(a, (int)(info & 0x1FFF));
            b = max(b, (int)((info >> 13) & 0x1FFF));
          }
        }
        rowInfo[h + bnd] = b >= 0 ? pack_row(a, b, false) : 0u;
      }
      const int xr0 = s_r0, xr1 = s_r1;
      __syncthreads();
      if (xr1 < 0)
        break; // empty intersection: both median finders give an empty image
      PH(1);
      // ---- histograms of bin = PixelType(10 * sqrtf(d2)) over the (i xor j) pixels, per side (X:413-459)
      // The pixels are dealt to the lanes across the WHOLE node, 32 consecutive pixels per warp visit: counts per
      // 32-word group -> prefix -> a visit looks its first group up and walks the groups its 32 pixels span.  Every
      // lane of a visit then has a pixel for the row search (the costly part), and every warp gets the same number
      // of visits, wherever the band of xor pixels happens to cross the node's rows.
      const int nGroups = (words + 31) >> 5;
      const bool dealt = nGroups <= TREE_GROUPS;
      if (dealt)
      {
        for (int g = warp; g < nGroups; g += TREE_WARPS)
        {
          const int wi = 32 * g + lane;
          int c = wi < words ? __popc(Sb[wi]) : 0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1)
            c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
          if (lane == 0)
            s_gcum[g + 1] = c;
        }
        __syncthreads();
        if (warp == 0)
        {
          // inclusive prefix over up to TREE_GROUPS counts, 32 at a time
          int carry = 0;
          for (int g0 = 0; g0 < nGroups; g0 += 32)
          {
            int v = g0 + lane < nGroups ? s_gcum[g0 + lane + 1] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
              const int u = __shfl_up_sync(0xFFFFFFFFu, v, o);
              if (lane >= o)
                v += u;
            }
            if (g0 + lane < nGroups)
              s_gcum[g0 + lane + 1] = carry + v;
            carry += __shfl_sync(0xFFFFFFFFu, v, 31);
          }
          if (lane == 0)
            s_gcum[0] = 0;
        }
        __syncthreads();
      }
      const int nPix = dealt ? s_gcum[nGroups] : 0;
      const int nVisits = dealt ? (nPix + 31) >> 5 : (words + TREE_HWORDS - 1) / TREE_HWORDS;
      for (int visit = warp; visit < nVisits; visit += TREE_WARPS)
      {
        // each lane ends up with: word index of its pixel (-1: none), bit, side
        int pIdx = -1, pBit = 0, pSide = 0;
        if (dealt)
        {
          const int p0 = 32 * visit, p = p0 + lane;
          int lo = 0, hi = nGroups; // first group g with s_gcum[g + 1] > p0
          while (lo < hi)
          {
            const int m = (lo + hi) >> 1;
            if (s_gcum[m + 1] > p0)
              hi = m;
            else
              lo = m + 1;
          }
          for (int g = lo; g < nGroups && s_gcum[g] < p0 + 32; ++g)
          {
            const int wi = 32 * g + lane;
            const uint32_t myS = wi < words ? Sb[wi] : 0u;
            const uint32_t myI = wi < words ? Ib[wi] : 0u;
            int incl = __popc(myS);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
              int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
              if (lane >= o)
                incl += v;
            }
            const int gBeg = s_gcum[g], gEnd = s_gcum[g + 1];
            const bool here = p >= gBeg && p < gEnd;
            int src, bit;
            warp_pixel(myS, incl, here ? p - gBeg : 0, src, bit);
            const uint32_t iw = __shfl_sync(0xFFFFFFFFu, myI, src);
            if (here)
            {
              pIdx = 32 * g + src;
              pBit = bit;
              pSide = (iw >> bit) & 1u ? 0 : 1;
            }
          }
        }
        else
        {
          // nodes too large for the group table: 8 words per visit, their pixels dealt 32 at a time below
          pIdx = -2;
        }
        const int wBeg = visit * TREE_HWORDS;
        uint32_t myS = 0, myI = 0;
        int incl = 0, total = dealt ? 32 : 0;
        if (!dealt)
        {
          const int myIdx = wBeg + lane;
          const bool mine = lane < TREE_HWORDS && myIdx < words;
          myS = mine ? Sb[myIdx] : 0u;
          myI = mine ? Ib[myIdx] : 0u;
          incl = __popc(myS);
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
          {
            int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o)
              incl += v;
          }
          total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
        for (int base = 0; base < total; base += 32)
        {
          bool on;
          int idx, bit, side;
          if (dealt)
          {
            on = pIdx >= 0;
            idx = pIdx;
            bit = pBit;
            side = pSide;
          }
          else
          {
            const int p = base + lane;
            on = p < total;
            int src;
            warp_pixel(myS, incl, on ? p : total - 1, src, bit);
            const uint32_t iw = __shfl_sync(0xFFFFFFFFu, myI, src);
            idx = wBeg + src;
            side = (iw >> bit) & 1u ? 0 : 1;
          }
          int d2 = 0;
          bool over = false;
          if (on)
          {
            const int r = idx / pitch, wi = idx - r * pitch;
            d2 = dt_edt_sq<false>(Xb, rowInfo, pitch, h, xr0, xr1, wi * 32 + bit, r, 0x7FFFFFFF);
            const int direct = dt_bin<T>(d2, P.err);
            if (direct < histSide)
            {
              atomicAdd(&hs[side * histSide + direct], 1u);
              atomicOr(&hsBm[direct >> 5], 1u << (direct & 31));
            }
            else
              over = true;
          }
          const unsigned needBig = __ballot_sync(0xFFFFFFFFu, over);
          if (needBig)
          {
            const int leader = __ffs(needBig) - 1;
            int bi = 0;
            if (lane == leader)
              bi = tree_get_big(&s_big, P.bigCount, P.nbig, P.err);
            bi = __shfl_sync(0xFFFFFFFFu, bi, leader);
            if (over)
            {
              const int bin = dt_bin<T>(d2, P.err);
              unsigned * bh = P.bigHist + (size_t)bi * DT_BIG_STRIDE;
              atomicAdd(&bh[side * 65536 + bin], 1u);
              atomicOr(&bh[2 * 65536 + (bin >> 5)], 1u << (bin & 31));
            }
          }
        }
      }
      __syncthreads();
      PH(2);
      // ---- scan: first bin minimising the score (X:461-494), threshold on d2 (X:497-508)
      {
        // a wrapped-bin table is in use for this node iff some bin reached it; the CTA keeps its table for later nodes,
        // so "in use" is decided by its bitmap: any touched bin
        unsigned * bh = s_big >= 0 ? P.bigHist + (size_t)s_big * DT_BIG_STRIDE : nullptr;
        int anyBig = 0;
        if (bh)
        {
          for (int k = tid; k < 2048; k += TREE_THREADS)
            anyBig |= bh[2 * 65536 + k] != 0u;
        }
        const bool big = __syncthreads_or(anyBig) != 0;
        unsigned * cntI = hs;
        unsigned * cntJ = hs + histSide;
        unsigned * bm = hsBm;
        if (big)
        {
          unsigned * bbm = bh + 2 * 65536;
          for (int wI = tid; wI < smallWords; wI += TREE_THREADS)
          {
            unsigned bits = hsBm[wI];
            if (!bits)
              continue;
            hsBm[wI] = 0u;
            while (bits)
            {
              int b = __ffs(bits) - 1;
              bits &= bits - 1;
              const int k = wI * 32 + b;
              atomicAdd(&bh[k], hs[k]);
              atomicAdd(&bh[65536 + k], hs[histSide + k]);
              atomicOr(&bbm[k >> 5], 1u << (k & 31));
              hs[k] = 0u;
              hs[histSide + k] = 0u;
            }
          }
          __syncthreads();
          cntI = bh;
          cntJ = bh + 65536;
          bm = bbm;
        }
        int thr = -1; // -1: no pixel outside the intersection, the median is the intersection (X:463-466)
        if (!big)
        {
          // the node's own table (the common case: distances of a few pixels, a few hundred bins): ONE warp does it,
          // no block-wide barrier, scan or reduction.  Bins are dealt to the lanes in contiguous ranges up to the
          // highest touched bin (lane order = bin order); untouched bins inside a range repeat the score of the bin
          // before them and lose every tie, so visiting them is exact; bin 0 is visited by lane 0 whether touched or
          // not (bestDiff starts at LLONG_MAX, X:482-494)
          if (warp == 0)
          {
            int mb = -1;
            for (int wI = lane; wI < smallWords; wI += 32)
            {
              const unsigned bits = hsBm[wI];
              if (bits)
              {
                mb = max(mb, wI * 32 + 31 - __clz(bits));
                hsBm[wI] = 0u; // leave the tables clean for the next node
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
              mb = max(mb, __shfl_xor_sync(0xFFFFFFFFu, mb, o));
            const int nb = mb + 1;
            const int per = (nb + 31) >> 5;
            const int k0 = min(nb, lane * per), k1 = min(nb, k0 + per);
            long long li = 0, lj = 0;
            for (int k = k0; k < k1; ++k)
            {
              li += cntI[k];
              lj += cntJ[k];
            }
            long long sa = li, sb = lj;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
              long long ta = __shfl_up_sync(0xFFFFFFFFu, sa, o), tb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
              if (lane >= o)
              {
                sa += ta;
                sb += tb;
              }
            }
            const long long iTot = __shfl_sync(0xFFFFFFFFu, sa, 31), jTot = __shfl_sync(0xFFFFFFFFu, sb, 31);
            if (iTot + jTot > 0)
            {
              long long si = sa - li, sj = sb - lj;
              long long bd = 0x7FFFFFFFFFFFFFFFll;
              int bidx = 0x7FFFFFFF;
              for (int k = k0; k < k1; ++k)
              {
                si += cntI[k];
                sj += cntJ[k];
                const long long d = llabs(llabs(iTot - si + sj) - llabs(jTot - sj + si));
                if (d < bd)
                {
                  bd = d;
                  bidx = k;
                }
                cntI[k] = 0u;
                cntJ[k] = 0u;
              }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1)
              {
                long long od = __shfl_xor_sync(0xFFFFFFFFu, bd, o);
                int oi = __shfl_xor_sync(0xFFFFFFFFu, bidx, o);
                if (od < bd || (od == bd && oi < bidx))
                {
                  bd = od;
                  bidx = oi;
                }
              }
              thr = dt_threshold_d2(bidx);
            }
          }
        }
        else
        {
          // bins are dealt to the threads in contiguous ranges up to the highest touched bin (thread order = bin order).
          // Untouched bins inside a range repeat the score of the bin before them and lose every tie, so visiting
          // them is exact; bin 0 is visited by thread 0 whether touched or not (bestDiff starts at LLONG_MAX, X:482-494)
          const int nWords = big ? 2048 : smallWords;
          if (tid == 0)
            s_maxBin = -1;
          __syncthreads();
          for (int wI = tid; wI < nWords; wI += TREE_THREADS)
          {
            const unsigned bits = bm[wI];
            if (bits)
            {
              atomicMax(&s_maxBin, wI * 32 + 31 - __clz(bits));
              bm[wI] = 0u; // leave the tables clean for the next node
            }
          }
          __syncthreads();
          const int nb = s_maxBin + 1;
          const int per = (nb + TREE_THREADS - 1) / TREE_THREADS;
          const int k0 = min(nb, tid * per), k1 = min(nb, k0 + per);
          long long li = 0, lj = 0;
          for (int k = k0; k < k1; ++k)
          {
            li += cntI[k];
            lj += cntJ[k];
          }
          long long sa = li, sb = lj;
  #pragma unroll
          for (int o = 1; o < 32; o <<= 1)
          {
            long long ta = __shfl_up_sync(0xFFFFFFFFu, sa, o), tb = __shfl_up_sync(0xFFFFFFFFu, sb, o);
            if (lane >= o)
            {
              sa += ta;
              sb += tb;
            }
          }
          if (lane == 31)
          {
            redA[warp] = sa;
            redB[warp] = sb;
          }
          __syncthreads();
          long long oa = 0, ob = 0, iTot = 0, jTot = 0;
          for (int k = 0; k < TREE_WARPS; ++k)
          {
            if (k < warp)
            {
              oa += redA[k];
              ob += redB[k];
            }
            iTot += redA[k];
            jTot += redB[k];
          }
          if (iTot + jTot > 0) // block-uniform
          {
            long long si = oa + sa - li, sj = ob + sb - lj;
            long long bd = 0x7FFFFFFFFFFFFFFFll;
            int bidx = 0x7FFFFFFF;
            for (int k = k0; k < k1; ++k)
            {
              si += cntI[k];
              sj += cntJ[k];
              const long long d = llabs(llabs(iTot - si + sj) - llabs(jTot - sj + si));
              if (d < bd)
              {
                bd = d;
                bidx = k;
              }
              cntI[k] = 0u;
              cntJ[k] = 0u;
            }
  #pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
              long long od = __shfl_xor_sync(0xFFFFFFFFu, bd, o);
              int oi = __shfl_xor_sync(0xFFFFFFFFu, bidx, o);
              if (od < bd || (od == bd && oi < bidx))
              {
                bd = od;
                bidx = oi;
              }
            }
            if (lane == 0)
            {
              bestD[warp] = bd;
              bestI[warp] = bidx;
            }
            __syncthreads();
            bd = bestD[0];
            bidx = bestI[0];
            for (int k = 1; k < TREE_WARPS; ++k)
              if (bestD[k] < bd || (bestD[k] == bd && bestI[k] < bidx))
              {
                bd = bestD[k];
                bidx = bestI[k];
              }
            thr = dt_threshold_d2(bidx);
          }
        }
        if (tid == 0)
          s_thr = thr;
      }
      __syncthreads();
      PH(3);
      // ---- median = intersection | (xor pixels with d2 <= threshold); bounding box inside the slice (X:497-515, 679-712)
      {
        const int thr = s_thr;
        const int axis = nd.axis;
        const int SU = (int)(axis == 0 ? P.VY : P.VX), SV = (int)(axis == 2 ? P.VY : P.VZ);
        int bx0 = 0x7FFFFFFF, bx1 = -0x7FFFFFFF, by0 = 0x7FFFFFFF, by1 = -0x7FFFFFFF;
        for (int myIdx = tid; myIdx < words; myIdx += TREE_THREADS)
        {
          const uint32_t myS = Sb[myIdx];
          uint32_t m = Xb[myIdx];
          const int r = myIdx / pitch, wi = myIdx - r * pitch;
          if (thr >= 0 && myS)
          {
            const int x0 = wi * 32;
            uint32_t pass = 0;
            int R = (int)sqrtf((float)thr) + 1;
            for (int k = 0; k * k <= thr; ++k)
            {
              while (R * R + k * k > thr)
                --R;
#pragma unroll
              for (int sgn = 0; sgn < 2; ++sgn)
              {
                if (sgn && k == 0)
                  continue;
                const int rr = sgn ? r + k : r - k;
                if (rr < xr0 || rr > xr1)
                  continue;
                const uint32_t info = rowInfo[rr];
                if (!info)
                  continue;
                const int a = info & 0x1FFF, b = (info >> 13) & 0x1FFF;
                if (a - R > x0 + 31 || b + R < x0)
                  continue;
                if ((info >> 26) & 1u)
                  pass |= bit_range(max(a - R, x0) - x0, min(b + R, x0 + 31) - x0);
                else
                {
                  const uint32_t * row = Xb + (size_t)rr * pitch;
                  const int wlo = max(0, (x0 - R) >> 5), whi = min(pitch - 1, (x0 + 31 + R) >> 5);
                  for (int ws = wlo; ws <= whi; ++ws)
                  {
                    uint32_t wv = row[ws];
                    while (wv)
                    {
                      const int sbit = __ffs(wv) - 1;
                      const uint32_t tt = ~(wv >> sbit);
                      const int len = tt ? __ffs(tt) - 1 : 32 - sbit;
                      const int lo = ws * 32 + sbit - R, hi = ws * 32 + sbit + len - 1 + R;
                      if (hi >= x0 && lo <= x0 + 31)
                        pass |= bit_range(max(lo, x0) - x0, min(hi, x0 + 31) - x0);
                      wv = (len >= 32) ? 0u : (wv & ~(((1u << len) - 1u) << sbit));
                    }
                  }
                }
              }
              if ((pass & myS) == myS)
                break;
            }
            m |= pass & myS;
          }
          Ib[myIdx] = m; // the median rows take the place of the i rows (the xor rows are still being read by others)
          const int y = uy0 + r;
          if (m && y >= 0 && y < SV)
          {
            const int xa = ux0 + 32 * wi;
            uint32_t c = m;
            if (xa < 0)
              c = (xa <= -32) ? 0u : (c & (0xFFFFFFFFu << (-xa)));
            if (xa + 32 > SU)
              c = (xa >= SU) ? 0u : (c & (0xFFFFFFFFu >> (xa + 32 - SU)));
            if (c)
            {
              bx0 = min(bx0, xa + __ffs(c) - 1);
              bx1 = max(bx1, xa + 31 - __clz(c));
              by0 = min(by0, y);
              by1 = max(by1, y);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          bx0 = min(bx0, __shfl_xor_sync(0xFFFFFFFFu, bx0, o));
          by0 = min(by0, __shfl_xor_sync(0xFFFFFFFFu, by0, o));
          bx1 = max(bx1, __shfl_xor_sync(0xFFFFFFFFu, bx1, o));
          by1 = max(by1, __shfl_xor_sync(0xFFFFFFFFu, by1, o));
        }
        if (lane == 0 && by1 >= by0)
        {
          atomicMin(&s_bx0, bx0);
          atomicMax(&s_bx1, bx1);
          atomicMin(&s_by0, by0);
          atomicMax(&s_by1, by1);
        }
      }
      __syncthreads();
      PH(4);
      // ---- mid shape: allocate, record, spawn the children (X:714-729, 766-792)
      if (tid == 0)
      {
        s_alive = 0;
        if (s_by1 >= s_by0)
        {
          MaskRef M;
          M.x0 = s_bx0;
          M.y0 = s_by0;
          M.w = s_bx1 - s_bx0 + 1;
          M.h = s_by1 - s_by0 + 1;
          M.pitch = (M.w + 31) >> 5;
          M.pad = 0;
          const int mwords = M.h * M.pitch;
          M.off = atomicAdd(P.arenaTop, (unsigned long long)mwords);
          if (M.off + (unsigned long long)mwords > P.arenaCap)
            atomicMax(P.err, ERR_ARENA);
          else
          {
            s_alive = 1;
            s_M = M;
            if (P.emit)
            {
              const int e = atomicAdd(P.emitCount, 1);
              if (e < P.emitCap)
              {
                EmitRec rec;
                rec.M = M;
                rec.M.off = M.off - P.emitBase;
                rec.axis = nd.axis;
                rec.label = nd.label;
                rec.mid = mid;
                rec.pad = 0;
                P.emit[e] = rec;
                if (nd.slot0 >= 0)
                  emit_slot_table(P.emit, P.emitCap)[nd.slot0 + mid - nd.lo - 1] = e;
              }
              else
                atomicMax(P.err, ERR_NODE_LIST);
            }
            if (abs(nd.i - nd.j) > 2)
            {
              const bool first = abs(nd.i - mid) > 1, second = abs(nd.j - mid) > 1;
              {
                if (second)
                {
                  Node c;
                  c.A = nd.B;
                  c.B = M;
                  c.tx = jT[0];
                  c.ty = jT[1];
                  c.i = nd.j;
                  c.j = mid;
                  c.axis = nd.axis;
                  c.label = nd.label;
                  c.slot0 = nd.slot0;
                  c.lo = nd.lo;
                  s_child[s_nchild++] = c;
                }
                if (first)
                {
                  Node c;
                  c.A = nd.A;
                  c.B = M;
                  c.tx = iT[0];
                  c.ty = iT[1];
                  c.i = nd.i;
                  c.j = mid;
                  c.axis = nd.axis;
                  c.label = nd.label;
                  c.slot0 = nd.slot0;
                  c.lo = nd.lo;
                  s_child[s_nchild++] = c;
                }
              }
            }
          }
        }
      }
      __syncthreads();
      if (!s_alive)
        break;
      PH(5);
      // ---- store the mid shape, publish the children, write the shape into the volume with the max rule (X:743-764)
      {
        constexpr int VPQ = 8 / sizeof(T);
        constexpr int GPW = 32 / VPQ + 1;
        const MaskRef M = s_M;
        const int axis = nd.axis;
        const T label = (T)nd.label;
        const unsigned long long nvox = (unsigned long long)P.VX * P.VY * P.VZ;
        long long su, sv;
        unsigned long long vbase;
        if (axis == 2)
        {
          su = 1;
          sv = P.VX;
          vbase = (unsigned long long)mid * P.VX * P.VY;
        }
        else if (axis == 1)
        {
          su = 1;
          sv = P.VX * P.VY;
          vbase = (unsigned long long)mid * P.VX;
        }
        else
        {
          su = P.VX;
          sv = P.VX * P.VY;
          vbase = (unsigned long long)mid;
        }
        const int mwords = M.h * M.pitch;
        // the mid shape goes to the arena first and the children are published right away: other CTAs start on them
        // while this one is still busy with the volume writes (read-modify-write round trips to L2 / HBM)
        for (int cell = tid; cell < mwords; cell += TREE_THREADS)
        {
          const int r = cell / M.pitch, wi = cell - r * M.pitch;
          const int y = M.y0 + r, x = M.x0 + 32 * wi;
          const uint32_t * row = Ib + (size_t)(y - uy0) * pitch;
          const int rx = x - ux0; // >= 0
          const int sw = rx >> 5, shf = rx & 31;
          const uint32_t lo = row[sw];
          const uint32_t hi = sw + 1 < pitch ? row[sw + 1] : 0u;
          uint32_t mw = __funnelshift_r(lo, hi, shf);
          const int valid = M.w - 32 * wi;
          if (valid < 32)
            mw &= (1u << valid) - 1u;
          P.arena[M.off + cell] = mw;
        }
        __threadfence();
        __syncthreads();
        if (tid == 0 && s_nchild)
        {
          const int n = s_nchild;
          s_nchild = 0;
          const int base = atomicAdd(P.rootNext + 2, n);
          if (base + n > P.queueCap)
            atomicMax(P.err, ERR_NODE_LIST);
          else
          {
            for (int k = 0; k < n; ++k)
              P.roots[nRoots + base + k] = s_child[k];
            __threadfence();
            for (int k = 0; k < n; ++k)
              vflags[base + k] = 1;
            atomicAdd(P.rootNext + 3, n);
          }
        }
        // slices perpendicular to z and to y: the compose passes after this kernel write them, no atomics
        const int mwordsW = (P.composeAxis2 && axis != 0) ? 0 : mwords;
        for (int wBeg = warp * 32; wBeg < mwordsW; wBeg += TREE_WARPS * 32)
        {
          const int wEnd = min(mwords, wBeg + 32);
          uint32_t myM = 0;
          {
            const int cell = wBeg + lane;
            if (cell < wEnd)
            {
              const int r = cell / M.pitch, wi = cell - r * M.pitch;
              const int y = M.y0 + r, x = M.x0 + 32 * wi;
              const uint32_t * row = Ib + (size_t)(y - uy0) * pitch;
              const int rx = x - ux0; // >= 0
              const int sw = rx >> 5, shf = rx & 31;
              const uint32_t lo = row[sw];
              const uint32_t hi = sw + 1 < pitch ? row[sw + 1] : 0u;
              myM = __funnelshift_r(lo, hi, shf);
              const int valid = M.w - 32 * wi;
              if (valid < 32)
                myM &= (1u << valid) - 1u;
            }
          }
          if (su == 1)
          {
            // x runs along the bits: aligned 8-byte groups; each lane keeps EMU groups in flight (loads first, then
            // the compare-and-swaps), the latency of a group's read-modify-write being what this phase costs
            constexpr int EMU = 3;
            const int nTasks = (wEnd - wBeg) * GPW;
            for (int tb = 0; tb < nTasks; tb += 32 * EMU)
            {
              unsigned long long g[EMU], qin[EMU], qout[EMU];
              unsigned bits[EMU];
#pragma unroll
              for (int u = 0; u < EMU; ++u)
              {
                const int task = tb + u * 32 + lane;
                const int c = min(task / GPW, wEnd - wBeg - 1), q = task % GPW;
                const uint32_t m = __shfl_sync(0xFFFFFFFFu, myM, c);
                bits[u] = 0u;
                g[u] = 0ull;
                if (task < nTasks && m)
                {
                  const int cell = wBeg + c;
                  const int r = cell / M.pitch, wi = cell - r * M.pitch;
                  const unsigned long long a0 = vbase + (unsigned long long)((long long)(M.x0 + 32 * wi) + (long long)(M.y0 + r) * sv);
                  g[u] = (a0 & ~(unsigned long long)(VPQ - 1)) + (unsigned long long)q * VPQ;
                  const long long sh = (long long)g[u] - (long long)a0; // in (-VPQ, 32]
                  if (sh >= 32)
                    bits[u] = 0u;
                  else if (sh >= 0)
                    bits[u] = (m >> sh) & ((1u << VPQ) - 1u);
                  else
                    bits[u] = (m << (-sh)) & ((1u << VPQ) - 1u);
                  if (bits[u] && g[u] + VPQ > nvox)
                  {
                    for (int k = 0; k < VPQ; ++k) // last, partial group of the volume
                      if (((bits[u] >> k) & 1u) && g[u] + k < nvox && in[g[u] + k] == 0)
                        atomic_max_pixel(out + g[u] + k, label);
                    bits[u] = 0u;
                  }
                }
                if (bits[u])
                {
                  qin[u] = __ldg(reinterpret_cast<const unsigned long long *>(in + g[u]));
                  qout[u] = *reinterpret_cast<const volatile unsigned long long *>(out + g[u]);
                }
              }
#pragma unroll
              for (int u = 0; u < EMU; ++u)
                if (bits[u])
                  write_group_loaded<T>(out, g[u], bits[u], label, qin[u], qout[u]);
            }
          }
          else if (nd.slot0 < 0) // registered axis-0 trees are written x-run by x-run afterwards (k_apply_axis0)
          {
            for (int c = 0; c < wEnd - wBeg; ++c)
            {
              const uint32_t m = __shfl_sync(0xFFFFFFFFu, myM, c);
              if ((m >> lane) & 1u)
              {
                const int cell = wBeg + c;
                const int r = cell / M.pitch, wi = cell - r * M.pitch;
                const unsigned long long a =
                  vbase + (unsigned long long)((long long)(M.x0 + 32 * wi + lane) * su + (long long)(M.y0 + r) * sv);
                if (in[a] == 0)
                  atomic_max_pixel(out + a, label);
              }
            }
          }
        }
      }
    } while (0);
    PH(6);
    // node finished (its children were published as soon as its shape was in the arena): count it as done -- after
    // the children, see the termination test above
    __syncthreads();
    if (tid == 0)
    {
      __threadfence();
      atomicAdd(P.rootNext + 4, 1);
    }
  }
#ifdef MCI_TREE_PROFILE
  if (tid == 0)
    for (int k = 0; k < 7; ++k)
      atomicAdd((unsigned long long *)(P.scratch + (size_t)gridDim.x * P.scratchPerCta) + k, (unsigned long long)phT[k]);
#endif
}

template <typename T>
static void
launch_dt_tree_t(Launch & L, int bit, const TreeParams & P, int grid, int threads, int smem)
{
  if (threads == 1024)
  {
    L.ensure_max_smem(k_dt_tree<T, 1024>, bit + 3, smem);
    launch_pdl(k_dt_tree<T, 1024>, grid, 1024, smem, L.stream, P);
  }
  else
  {
    L.ensure_max_smem(k_dt_tree<T, 512>, bit, smem);
    launch_pdl(k_dt_tree<T, 512>, grid, 512, smem, L.stream, P);
  }
}

void
launch_dt_tree(Launch & L, int ptype, const TreeParams & P, int nRootsMax, int grid, int threads)
{
  if (nRootsMax <= 0)
    return;
  const int smem = (2 * P.histSide + P.histSide / 32 + TREE_NODE_WORDS) * 4;
  L.pre("k_dt_tree");
  switch (ptype)
  {
    case MCI_U8:
      launch_dt_tree_t<uint8_t>(L, 8, P, grid, threads, smem);
      break;
    case MCI_U16:
      launch_dt_tree_t<uint16_t>(L, 9, P, grid, threads, smem);
      break;
    default:
      launch_dt_tree_t<int16_t>(L, 10, P, grid, threads, smem);
  }
  L.post("k_dt_tree");
}

// ===================================================================================================
// Compose pass for the mid slices perpendicular to z (X:743-764 with X:1623-1636 folded in):
// instead of every node pushing its shape into the volume with compare-and-swap groups (three dependent round trips
// to L2 per 8 bytes, the largest single cost of k_dt_tree), the shapes are binned by z-plane and ONE thread per voxel
// takes the largest label among the shapes covering it: out = max(out, label) where in = 0 -- plain coalesced loads
// and stores, no atomics, nobody else writes the voxel during the pass (slices along x and y are written by other
// kernels, stream-ordered after this one).  The max rule makes the order of the shapes irrelevant.
//   k_compose_count : per plane, number of shapes and their common bounding box
//   k_compose_scan  : exclusive prefix over the planes
//   k_compose_fill  : record indices grouped by plane
//   k_compose       : one warp per row of a plane's box, 32 pixels per step
// ===================================================================================================
struct ComposeTables
{
  int * count; // [VZ]
  int * start; // [VZ + 1]
  int * fill;  // [VZ]
  int * box;   // [4 * VZ]  min x, min y, max x, max y
  int * recs;  // [emitCap]
};

__global__ void __launch_bounds__(256)
k_compose_count(const EmitRec * __restrict__ recs, const int * __restrict__ nrecsDev, int cap, ComposeTables t, int VZ, int axis)
{
  pdl_wait();
  const int n = min(*nrecsDev, cap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
  {
    const EmitRec r = recs[e];
    if (r.axis != axis || (unsigned)r.mid >= (unsigned)VZ)
      continue;
    atomicAdd(&t.count[r.mid], 1);
    atomicMin(&t.box[4 * r.mid + 0], r.M.x0);
    atomicMin(&t.box[4 * r.mid + 1], r.M.y0);
    atomicMax(&t.box[4 * r.mid + 2], r.M.x0 + r.M.w - 1);
    atomicMax(&t.box[4 * r.mid + 3], r.M.y0 + r.M.h - 1);
  }
}

__global__ void __launch_bounds__(1024)
k_compose_scan(ComposeTables t, int VZ)
{
  pdl_wait();
  __shared__ int warpSum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
    carry = 0;
  __syncthreads();
  for (int base = 0; base < VZ; base += 1024)
  {
    const int z = base + tid;
    const int c = z < VZ ? t.count[z] : 0;
    int v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int u = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o)
        v += u;
    }
    if (lane == 31)
      warpSum[warp] = v;
    __syncthreads();
    int off = carry;
    for (int k = 0; k < warp; ++k)
      off += warpSum[k];
    if (z < VZ)
      t.start[z] = off + v - c;
    __syncthreads();
    if (tid == 1023)
      carry = off + v;
    __syncthreads();
  }
  if (tid == 0)
    t.start[VZ] = carry;
}

__global__ void __launch_bounds__(256)
k_compose_fill(const EmitRec * __restrict__ recs, const int * __restrict__ nrecsDev, int cap, ComposeTables t, int VZ, int axis)
{
  pdl_wait();
  const int n = min(*nrecsDev, cap);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
  {
    const EmitRec r = recs[e];
    if (r.axis != axis || (unsigned)r.mid >= (unsigned)VZ)
      continue;
    t.recs[t.start[r.mid] + atomicAdd(&t.fill[r.mid], 1)] = e;
  }
}

// The three table kernels and the three memsets before them in ONE single-CTA launch, for runs with few mid slices (a
// C3 step has 938): initialise, count + boxes, prefix, record lists, with CTA barriers in between.  The counters are
// read back with L2 loads (the atomics that raised them were performed at L2).
#define COMPOSE_SMALL_CAP 8192
__global__ void __launch_bounds__(1024)
k_compose_tables(const EmitRec * __restrict__ recs, const int * __restrict__ nrecsDev, int cap, ComposeTables t, int VZ, int axis)
{
  pdl_wait();
  __shared__ int warpSum[32];
  __shared__ int carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int z = tid; z < VZ; z += 1024)
  {
    t.count[z] = 0;
    t.fill[z] = 0;
    t.box[4 * z + 0] = 0x7F7F7F7F;
    t.box[4 * z + 1] = 0x7F7F7F7F;
    t.box[4 * z + 2] = (int)0x80808080;
    t.box[4 * z + 3] = (int)0x80808080;
  }
  if (tid == 0)
    carry = 0;
  __syncthreads();
  const int n = min(*nrecsDev, cap);
  for (int e = tid; e < n; e += 1024)
  {
    const EmitRec r = recs[e];
    if (r.axis != axis || (unsigned)r.mid >= (unsigned)VZ)
      continue;
    atomicAdd(&t.count[r.mid], 1);
    atomicMin(&t.box[4 * r.mid + 0], r.M.x0);
    atomicMin(&t.box[4 * r.mid + 1], r.M.y0);
    atomicMax(&t.box[4 * r.mid + 2], r.M.x0 + r.M.w - 1);
    atomicMax(&t.box[4 * r.mid + 3], r.M.y0 + r.M.h - 1);
  }
  __threadfence();
  __syncthreads();
  for (int base = 0; base < VZ; base += 1024)
  {
    const int z = base + tid;
    const int c = z < VZ ? __ldcg(&t.count[z]) : 0;
    int v = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int u = __shfl_up_sync(0xFFFFFFFFu, v, o);
      if (lane >= o)
        v += u;
    }
    if (lane == 31)
      warpSum[warp] = v;
    __syncthreads();
    int off = carry;
    for (int k = 0; k < warp; ++k)
      off += warpSum[k];
    if (z < VZ)
      t.start[z] = off + v - c;
    __syncthreads();
    if (tid == 1023)
      carry = off + v;
    __syncthreads();
  }
  if (tid == 0)
    t.start[VZ] = carry;
  __threadfence();
  __syncthreads();
  for (int e = tid; e < n; e += 1024)
  {
    const EmitRec r = recs[e];
    if (r.axis != axis || (unsigned)r.mid >= (unsigned)VZ)
      continue;
    t.recs[__ldcg(&t.start[r.mid]) + atomicAdd(&t.fill[r.mid], 1)] = e;
  }
}

#define COMPOSE_ROWS 8  // rows of a plane per CTA (one warp each)
#define COMPOSE_LIST 64 // shapes crossing a row kept per pass (more: further passes over the row)
struct ComposeEntry
{
  const uint32_t * row; // the shape's bit row at this y
  int x0, w, pitch, label;
};
// Measured alternatives (C3 / C5, ms): this version 0.052 / 1.48; one pixel per lane instead of 8-byte groups 0.051 /
// 1.54; the plane's shapes staged once per CTA, 16 rows per CTA 0.054 / 1.78; x-major (largest label per voxel over all
// shapes crossing the row first, then one read-modify-write per group) 0.090 / 2.75 -- the shapes crossing a row lie
// far apart, so sweeping their common extent is mostly empty steps.
template <typename T>
__global__ void __launch_bounds__(32 * COMPOSE_ROWS)
k_compose(const EmitRec * __restrict__ recs, const uint32_t * __restrict__ words, ComposeTables t, const T * __restrict__ in, T * out,
          long long VX, long long VY, int axis)
{
  pdl_wait();
  __shared__ ComposeEntry s_list[COMPOSE_ROWS][COMPOSE_LIST];
  const int z = blockIdx.y;
  const int n = t.count[z];
  if (n == 0)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int by0 = t.box[4 * z + 1], by1 = t.box[4 * z + 3];
  const int y = by0 + (int)blockIdx.x * COMPOSE_ROWS + warp;
  if (y > by1)
    return; // no block-wide barrier below
  const int * list = t.recs + t.start[z];
  ComposeEntry * L = s_list[warp];
  // slices perpendicular to z (axis 2): plane = z, row = y; perpendicular to y (axis 1): plane = y, row = z (X:681-693);
  // either way a shape's bits run along x, i.e. along contiguous voxels
  const unsigned long long rowBase = axis == 2
                                       ? ((unsigned long long)z * (unsigned long long)VY + (unsigned long long)y) * (unsigned long long)VX
                                       : ((unsigned long long)y * (unsigned long long)VY + (unsigned long long)z) * (unsigned long long)VX;
  constexpr int VPQ = 8 / sizeof(T);
  const bool vec = (VX % VPQ) == 0;
  for (int first = 0; first < n;)
  {
    // the shapes of this plane that cross row y, 32 candidates at a time while a whole batch still fits
    int kept = 0, e = first;
    while (e < n && kept + 32 <= COMPOSE_LIST)
    {
      const int idx = e + lane;
      bool hit = false;
      EmitRec r;
      if (idx < n)
      {
        r = recs[list[idx]];
        hit = y >= r.M.y0 && y < r.M.y0 + r.M.h;
      }
      const unsigned vote = __ballot_sync(0xFFFFFFFFu, hit);
      if (hit)
      {
        ComposeEntry c;
        c.row = words + r.M.off + (unsigned long long)(y - r.M.y0) * (unsigned long long)r.M.pitch;
        c.x0 = r.M.x0;
        c.w = r.M.w;
        c.pitch = r.M.pitch;
        c.label = r.label;
        L[kept + __popc(vote & ((1u << lane) - 1u))] = c;
      }
      kept += __popc(vote);
      e += 32;
    }
    first = e;
    __syncwarp();
    // shape after shape: this warp is the only writer of the row, so read-modify-write needs no atomics; the
    // __syncwarp between shapes orders one shape's stores before the next one's loads of the same voxels.
    // Rows whose voxels are 8-byte aligned take VPQ voxels per lane (one 8-byte load of in and out, one store): a warp
    // covers 32 * VPQ pixels per step, the shape's bits aligned to that grid with a funnel shift.
    for (int k = 0; k < kept; ++k)
    {
      const ComposeEntry c = L[k]; // broadcast read
      const T lab = (T)c.label;
      if (vec)
      {
        union Q
        {
          unsigned long long u;
          T v[VPQ];
        };
        const int span = 32 * VPQ;
        for (int X0 = c.x0 & ~(span - 1); X0 < c.x0 + c.w; X0 += span)
        {
          const int x = X0 + VPQ * lane;        // first voxel of this lane's group
          const int rx = x - c.x0;              // bit position in the shape's row (may be negative / beyond w)
          const int wi = rx >> 5, sh = rx & 31;
          const uint32_t lo = ((unsigned)wi < (unsigned)c.pitch) ? __ldg(c.row + wi) : 0u;
          const uint32_t hi = ((unsigned)(wi + 1) < (unsigned)c.pitch) ? __ldg(c.row + wi + 1) : 0u;
          const unsigned bits = __funnelshift_r(lo, hi, sh) & ((1u << VPQ) - 1u);
          if (bits)
          {
            const unsigned long long a = rowBase + (unsigned long long)x; // x >= 0: bits are zero left of the shape
            Q qi, qo, qn;
            qi.u = __ldg(reinterpret_cast<const unsigned long long *>(in + a));
            qo.u = *reinterpret_cast<const unsigned long long *>(out + a);
            qn = qo;
#pragma unroll
            for (int q = 0; q < VPQ; ++q)
              if (((bits >> q) & 1u) && qi.v[q] == 0 && qo.v[q] < lab)
                qn.v[q] = lab;
            if (qn.u != qo.u)
              *reinterpret_cast<unsigned long long *>(out + a) = qn.u;
          }
        }
      }
      else
      {
        for (int w0 = 0; w0 < c.pitch; ++w0)
        {
          const uint32_t m = __ldg(c.row + w0);
          if ((m >> lane) & 1u)
          {
            const unsigned long long a = rowBase + (unsigned long long)(c.x0 + 32 * w0 + lane);
            if (in[a] == 0 && out[a] < lab)
              out[a] = lab;
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();
  }
}

// few records: k_compose_tables initialises the tables itself (the caller then skips its memsets)
bool
compose_tables_in_one_launch(int cap)
{
  static const bool off = std::getenv("MCI_COMPOSE_3K") != nullptr; // A/B switch: the three table kernels
  return !off && cap <= COMPOSE_SMALL_CAP;
}

void
launch_compose(Launch & L, int ptype, int axis, const EmitRec * recs, const int * nrecsDev, int cap, const uint32_t * words, int * tables,
               const void * in, void * out, const long long dims[3])
{
  if (cap <= 0 || (axis != 1 && axis != 2))
    return;
  const int nPlanes = (int)dims[axis];                 // slice positions along the axis
  const long long nRows = axis == 2 ? dims[1] : dims[2]; // second coordinate of a slice
  ComposeTables t;
  t.count = tables;
  t.start = t.count + nPlanes;
  t.fill = t.start + nPlanes + 1;
  t.box = t.fill + nPlanes;
  t.recs = t.box + 4 * nPlanes;
  const int g = std::max(1, std::min((cap + 255) / 256, L.sms * 4));
  if (compose_tables_in_one_launch(cap))
  {
    L.pre("k_compose_tables");
    launch_pdl(k_compose_tables, 1, 1024, 0, L.stream, recs, nrecsDev, cap, t, nPlanes, axis);
    L.post("k_compose_tables");
  }
  else
  {
  L.pre("k_compose_count");
  launch_pdl(k_compose_count, g, 256, 0, L.stream, recs, nrecsDev, cap, t, nPlanes, axis);
  L.post("k_compose_count");
  L.pre("k_compose_scan");
  launch_pdl(k_compose_scan, 1, 1024, 0, L.stream, t, nPlanes);
  L.post("k_compose_scan");
  L.pre("k_compose_fill");
  launch_pdl(k_compose_fill, g, 256, 0, L.stream, recs, nrecsDev, cap, t, nPlanes, axis);
  L.post("k_compose_fill");
  }
  const dim3 grid((unsigned)((nRows + COMPOSE_ROWS - 1) / COMPOSE_ROWS), (unsigned)nPlanes);
  L.pre("k_compose");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_compose<uint8_t>, grid, 32 * COMPOSE_ROWS, 0, L.stream, recs, words, t, (const uint8_t *)in, (uint8_t *)out, dims[0], dims[1], axis);
      break;
    case MCI_U16:
      launch_pdl(k_compose<uint16_t>, grid, 32 * COMPOSE_ROWS, 0, L.stream, recs, words, t, (const uint16_t *)in, (uint16_t *)out, dims[0], dims[1], axis);
      break;
    default:
      launch_pdl(k_compose<int16_t>, grid, 32 * COMPOSE_ROWS, 0, L.stream, recs, words, t, (const int16_t *)in, (int16_t *)out, dims[0], dims[1], axis);
  }
  L.post("k_compose");
}

void
launch_dt_level(Launch & L, int ptype, const DtLevel & lv)
{
  if (lv.maxNodes <= 0)
    return;
  const int nodeGrid = std::max(1, std::min(lv.maxNodes, L.sms * 8));
  const int itemGrid = std::max(1, std::min(lv.itemCap, L.sms * 8));
  const int histGrid = std::max(1, std::min((lv.hitemCap + 7) / 8, L.sms * 8));
  L.pre("k_dt_prep");
  launch_pdl(k_dt_prep, nodeGrid, 256, 0, L.stream, lv.nodes, lv.count, lv.work, lv.arena, lv.scratch, lv.scratchTop, lv.scratchCap,
             lv.items, lv.itemCount, lv.itemCap, lv.hitems, lv.hitemCount, lv.hitemCap, lv.err);
  L.post("k_dt_prep");
  L.pre("k_dt_hist");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_dt_hist<uint8_t>, histGrid, 256, 0, L.stream, lv.work, lv.hitems, lv.hitemCount, lv.hitemCap, lv.scratch, lv.hist, lv.histStride,
                                                          lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
      break;
    case MCI_U16:
      launch_pdl(k_dt_hist<uint16_t>, histGrid, 256, 0, L.stream, lv.work, lv.hitems, lv.hitemCount, lv.hitemCap, lv.scratch, lv.hist, lv.histStride,
                                                           lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
      break;
    default:
      launch_pdl(k_dt_hist<int16_t>, histGrid, 256, 0, L.stream, lv.work, lv.hitems, lv.hitemCount, lv.hitemCap, lv.scratch, lv.hist, lv.histStride,
                                                          lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
  }
  L.post("k_dt_hist");
  L.pre("k_dt_scan");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_dt_scan<uint8_t>, nodeGrid, 256, 0, L.stream, lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
      break;
    case MCI_U16:
      launch_pdl(k_dt_scan<uint16_t>, nodeGrid, 256, 0, L.stream, lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
      break;
    default:
      launch_pdl(k_dt_scan<int16_t>, nodeGrid, 256, 0, L.stream, lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
  }
  L.post("k_dt_scan");
  L.pre("k_dt_median");
  launch_pdl(k_dt_median, itemGrid, 256, 0, L.stream, lv.work, lv.items, lv.itemCount, lv.scratch, (int)lv.dims[0], (int)lv.dims[1],
                                               (int)lv.dims[2], lv.err);
  L.post("k_dt_median");
  L.pre("k_dt_alloc");
  launch_pdl(k_dt_alloc, (lv.maxNodes + 127) / 128, 128, 0, L.stream, lv.work, lv.count, lv.arenaTop, lv.arenaCap, lv.next, lv.nextCount,
                                                                lv.nextCap, lv.witems, lv.witemCount, lv.itemCap, lv.emit, lv.emitCount,
                                                                lv.emitCap, lv.emitBase, lv.err);
  L.post("k_dt_alloc");
  L.pre("k_dt_emit");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_dt_emit<uint8_t>, itemGrid, 256, 0, L.stream, lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                          (const uint8_t *)lv.in, (uint8_t *)lv.out, lv.dims[0], lv.dims[1],
                                                          lv.dims[2]);
      break;
    case MCI_U16:
      launch_pdl(k_dt_emit<uint16_t>, itemGrid, 256, 0, L.stream, lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                           (const uint16_t *)lv.in, (uint16_t *)lv.out, lv.dims[0], lv.dims[1],
                                                           lv.dims[2]);
      break;
    default:
      launch_pdl(k_dt_emit<int16_t>, itemGrid, 256, 0, L.stream, lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                          (const int16_t *)lv.in, (int16_t *)lv.out, lv.dims[0], lv.dims[1],
                                                          lv.dims[2]);
  }
  L.post("k_dt_emit");
}

} // namespace mci
