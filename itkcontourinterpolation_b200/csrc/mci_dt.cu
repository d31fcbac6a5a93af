// mci_dt.cu -- one level of the Interpolate1to1 recursion with the distance-transform median
// (reference X:556-793 with FindMedianImageDistances X:405-516), as six grid-wide kernels:
//
//   k_dt_prep    per node : translation halving, union region, translate+binarise+AND into bit rows,
//                           per-row extents of the intersection, work items            (X:570-666)
//   k_dt_hist    per item : exact squared distance of every (i xor j) pixel to the intersection,
//                           histograms per side                                          (X:413-459)
//   k_dt_scan    per node : prefix sums, first bin minimising the score, threshold       (X:461-494)
//   k_dt_median  per item : median = intersection | (xor pixels with d2 <= threshold), bounding box (X:497-515)
//   k_dt_alloc   per node : crop/clip, allocate the mid shape, spawn the children        (X:679-729, 766-792)
//   k_dt_emit    per item : store the mid shape, write it into the volume with the max rule (X:743-764)
//
// Work items are fixed-size word ranges of a node's bit rows, so the whole GPU works on a level no
// matter how few nodes it has or how unequal they are.
#include <algorithm>
#include <type_traits>
#include "mci_kernels.cuh"
#include "mci_launch.h"

namespace mci
{

#define DT_CHUNK 256 // words of a node's bit rows per work item (one CTA: 8 warps x 32 words)

__device__ __forceinline__ uint32_t
pack_row(int a, int b, bool single)
{
  return (uint32_t)a | ((uint32_t)b << 13) | (single ? (1u << 26) : 0u) | (1u << 27);
}

// exact squared distance from pixel (x, r) to the nearest intersection pixel, or `cap` if not below cap.
// rowInfo[r] packs the first / last set column of row r and whether the row is one run (then the
// distance to the row is the distance to that interval); 0 = empty row.
__device__ __forceinline__ int
dt_edt_sq(const uint32_t * __restrict__ Xb, const uint32_t * __restrict__ rowInfo, int pitch, int xr0, int xr1, int x, int r,
          int cap)
{
  int best = cap;
  int k = max(0, max(xr0 - r, r - xr1));
  for (;; ++k)
  {
    const int kk = k * k;
    if (kk >= best)
      break;
    const int r1 = r - k, r2 = r + k;
    if (r1 < xr0 && r2 > xr1)
      break;
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
      const int rr = s ? r2 : r1;
      if (rr < xr0 || rr > xr1 || (s && k == 0))
        continue;
      const uint32_t info = __ldg(rowInfo + rr);
      if (!info)
        continue;
      const int a = info & 0x1FFF, b = (info >> 13) & 0x1FFF;
      const int dx = x < a ? a - x : (x > b ? x - b : 0);
      const int lo = kk + dx * dx;
      if (lo >= best)
        continue;
      if ((info >> 26) & 1u)
        best = lo;
      else
        best = min(best, kk + row_nearest_sq(Xb + (size_t)rr * pitch, pitch, x, best - kk));
    }
  }
  return best;
}

template <typename T>
__device__ __forceinline__ int
dt_bin(int d2, int * err)
{
  int q = __float2int_rz(__fmul_rn(10.0f, (float)d2)); // `PixelType dist = fractioning * sdf` (X:431)
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

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dt_prep(const Node * __restrict__ nodes, const int * __restrict__ count, NodeWork * work, const uint32_t * __restrict__ arena,
          uint32_t * scratch, unsigned long long * scratchTop, unsigned long long scratchCap, DtItem * items, int * itemCount,
          int itemCap, int * err)
{
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
      s_off = atomicAdd(scratchTop, (unsigned long long)(3 * (size_t)words + h));
    }
    __syncthreads();
    const unsigned long long off = s_off;
    NodeWork * W = work + nidx;
    if (off + 3ull * words + h > scratchCap || w > 8191)
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
    // translate, binarise, intersect (X:610-666); one warp per row so that row extents come for free
    int r0 = 0x7FFFFFFF, r1 = -1;
    for (int r = warp; r < h; r += 8)
    {
      const int y = uy0 + r;
      int first = 0x7FFFFFFF, last = -1, runs = 0;
      uint32_t prevTop = 0;
      for (int wb = 0; wb < pitch; wb += 32)
      {
        const int wi = wb + lane;
        uint32_t a = 0, b = 0;
        if (wi < pitch)
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
        uint32_t below = __shfl_up_sync(0xFFFFFFFFu, xw, 1);
        if (lane == 0)
          below = prevTop;
        const uint32_t starts = xw & ~((xw << 1) | (below >> 31));
        runs += __popc(starts);
        if (xw)
        {
          first = min(first, 32 * wi + __ffs(xw) - 1);
          last = max(last, 32 * wi + 31 - __clz(xw));
        }
        prevTop = __shfl_sync(0xFFFFFFFFu, xw, 31);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        first = min(first, __shfl_xor_sync(0xFFFFFFFFu, first, o));
        last = max(last, __shfl_xor_sync(0xFFFFFFFFu, last, o));
        runs += __shfl_xor_sync(0xFFFFFFFFu, runs, o);
      }
      if (lane == 0)
      {
        rowInfo[r] = last >= 0 ? pack_row(first, last, runs == 1) : 0u;
        if (last >= 0)
        {
          r0 = min(r0, r);
          r1 = max(r1, r);
        }
      }
    }
    if (lane == 0 && r1 >= 0)
    {
      atomicMin(&s_r0, r0);
      atomicMax(&s_r1, r1);
    }
    __syncthreads();
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
      nw.maxIdx = -1;
      nw.maxBig = -1;
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

template <typename T>
__global__ void __launch_bounds__(256)
k_dt_hist(NodeWork * work, const DtItem * __restrict__ items, const int * __restrict__ itemCount, const uint32_t * __restrict__ scratch,
          unsigned * hist, int histStride, int histSide, unsigned * bigHist, int * bigCount, int nbig, int * err)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nItems = *itemCount;
  constexpr bool kBinsOnly = sizeof(T) == 1; // uchar: 256 wrapped bins, indexed directly
  constexpr int kSmall = PixTraits<T>::kNoWrapD2;
  for (int it = blockIdx.x; it < nItems; it += gridDim.x)
  {
    const DtItem item = items[it];
    NodeWork * W = work + item.node;
    const int pitch = W->pitch, words = W->words, xr0 = W->xr0, xr1 = W->xr1;
    const uint32_t * Xb = scratch + W->scratchOff;
    const uint32_t * Ib = Xb + words;
    const uint32_t * Sb = Ib + words;
    const uint32_t * rowInfo = Sb + words;
    unsigned * h0 = hist + (size_t)item.node * histStride;
    int maxSmall = -1, maxBig = -1;
    const int wBeg = item.w0 + warp * 32, wEnd = min(words, wBeg + 32);
    for (int idx = wBeg; idx < wEnd; ++idx)
    {
      const uint32_t s = __ldg(Sb + idx);
      if (!s)
        continue; // warp-uniform
      const bool on = (s >> lane) & 1u;
      int d2 = 0, side = 0;
      if (on)
      {
        const int r = idx / pitch, wi = idx - r * pitch;
        d2 = dt_edt_sq(Xb, rowInfo, pitch, xr0, xr1, wi * 32 + lane, r, 0x7FFFFFFF);
        side = (__ldg(Ib + idx) >> lane) & 1u ? 0 : 1;
      }
      bool over = false;
      if (on)
      {
        if (kBinsOnly)
        {
          int bin = dt_bin<T>(d2, err);
          atomicAdd(&h0[side * histSide + bin], 1u);
          maxSmall = max(maxSmall, bin);
        }
        else if (d2 <= kSmall)
        {
          atomicAdd(&h0[side * histSide + d2], 1u);
          maxSmall = max(maxSmall, d2);
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
          int bin = dt_bin<T>(d2, err);
          atomicAdd(&bigHist[(size_t)bi * 2 * 65536 + side * 65536 + bin], 1u);
          maxBig = max(maxBig, bin);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      maxSmall = max(maxSmall, __shfl_xor_sync(0xFFFFFFFFu, maxSmall, o));
      maxBig = max(maxBig, __shfl_xor_sync(0xFFFFFFFFu, maxBig, o));
    }
    if (lane == 0)
    {
      if (maxSmall >= 0)
        atomicMax(&W->maxIdx, maxSmall);
      if (maxBig >= 0)
        atomicMax(&W->maxBig, maxBig);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_dt_scan(NodeWork * work, const int * __restrict__ count, unsigned * hist, int histStride, int histSide, unsigned * bigHist)
{
  __shared__ long long redA[8], redB[8], bestD[8];
  __shared__ int bestI[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = *count;
  constexpr bool kBinsOnly = sizeof(T) == 1;
  for (int nidx = blockIdx.x; nidx < n; nidx += gridDim.x)
  {
    NodeWork * W = work + nidx;
    if (!W->alive)
      continue;
    unsigned * hs = hist + (size_t)nidx * histStride;
    const int maxSmall = W->maxIdx, maxBig = W->maxBig, bi = W->bigIdx;
    const bool big = bi >= 0;
    unsigned * hi = hs;
    unsigned * hj = hs + histSide;
    int len; // histogram length (maxSize, X:462), in the index domain being scanned
    __syncthreads();
    if (big)
    {
      // some distance wrapped (or exceeded the direct table): work in the wrapped-bin domain of X:431.
      // fold the directly indexed part in (bin = 10*d2 below the wrap)
      unsigned * bh = bigHist + (size_t)bi * 2 * 65536;
      int top = maxBig;
      for (int k = tid; k <= maxSmall; k += 256)
      {
        unsigned ci = hs[k], cj = hs[histSide + k];
        if (ci)
          atomicAdd(&bh[10 * k], ci);
        if (cj)
          atomicAdd(&bh[65536 + 10 * k], cj);
        hs[k] = 0u;
        hs[histSide + k] = 0u;
      }
      if (maxSmall >= 0)
        top = max(top, 10 * maxSmall);
      len = top + 1;
      hi = bh;
      hj = bh + 65536;
      __syncthreads();
    }
    else
      len = maxSmall + 1;
    int thr = -1; // -1: no pixel outside the intersection, the median is the intersection (X:463-466)
    if (len > 0)
    {
      const int chunk = (len + 255) / 256;
      const int lo = min(len, tid * chunk), hiEnd = min(len, lo + chunk);
      long long ci = 0, cj = 0;
      for (int k = lo; k < hiEnd; ++k)
      {
        ci += hi[k];
        cj += hj[k];
      }
      long long sa = ci, sb = cj;
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
      long long si = oa + sa - ci, sj = ob + sb - cj;
      long long bd = 0x7FFFFFFFFFFFFFFFll;
      int bidx = 0x7FFFFFFF;
      for (int k = lo; k < hiEnd; ++k)
      {
        si += hi[k];
        sj += hj[k];
        long long a = llabs(iTot - si + sj), b = llabs(jTot - sj + si);
        long long d = llabs(a - b);
        if (d < bd)
        {
          bd = d;
          bidx = k;
        }
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
      // threshold `sdf <= float(bestBin) / 10` on integer d2 (X:497-508)
      if (big || kBinsOnly)
        thr = (int)floorf(__fdiv_rn((float)bidx, 10.0f));
      else
        thr = bidx; // bestBin = 10 * d2
      __syncthreads();
      // leave the tables clean for the next user
      for (int k = tid; k < len; k += 256)
      {
        hi[k] = 0u;
        hj[k] = 0u;
      }
    }
    if (tid == 0)
      W->thr = thr;
  }
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dt_median(NodeWork * work, const DtItem * __restrict__ items, const int * __restrict__ itemCount, uint32_t * scratch, int SU0,
            int SU1, int SU2, int * err)
{
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
    int bx0 = 0x7FFFFFFF, bx1 = -0x7FFFFFFF, by0 = 0x7FFFFFFF, by1 = -0x7FFFFFFF;
    const int wBeg = item.w0 + warp * 32, wEnd = min(words, wBeg + 32);
    for (int idx = wBeg; idx < wEnd; ++idx)
    {
      const uint32_t s = Sb[idx];
      const uint32_t x = __ldg(Xb + idx);
      if (!(s | x))
        continue; // warp-uniform
      const int r = idx / pitch, wi = idx - r * pitch;
      uint32_t m = x;
      if (s && thr >= 0)
      {
        bool pass = false;
        if ((s >> lane) & 1u)
          pass = dt_edt_sq(Xb, rowInfo, pitch, xr0, xr1, wi * 32 + lane, r, thr + 1) <= thr;
        m |= __ballot_sync(0xFFFFFFFFu, pass);
      }
      if (lane == 0)
        Sb[idx] = m;
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
          bx0 = min(bx0, xa + __ffs(c) - 1);
          bx1 = max(bx1, xa + 31 - __clz(c));
          by0 = min(by0, y);
          by1 = max(by1, y);
        }
      }
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
           int * nextCount, int nextCap, DtItem * witems, int * witemCount, int witemCap, int * err)
{
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
        next[pos++] = c;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_dt_emit(const NodeWork * __restrict__ work, const DtItem * __restrict__ witems, const int * __restrict__ witemCount,
          const uint32_t * __restrict__ scratch, uint32_t * arena, const T * __restrict__ in, T * out, long long VX, long long VY,
          long long VZ)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nItems = *witemCount;
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
    for (int cell = wBeg; cell < wEnd; ++cell)
    {
      const int r = cell / M.pitch, wi = cell - r * M.pitch;
      const int y = M.y0 + r, x = M.x0 + 32 * wi;
      const uint32_t * row = Sb + (size_t)(y - uy0) * pitch;
      const int rx = x - ux0; // >= 0
      const int sw = rx >> 5, shf = rx & 31;
      const uint32_t lo = __ldg(row + sw);
      const uint32_t hi = sw + 1 < pitch ? __ldg(row + sw + 1) : 0u;
      uint32_t m = __funnelshift_r(lo, hi, shf);
      const int valid = M.w - 32 * wi;
      if (valid < 32)
        m &= (1u << valid) - 1u;
      if (lane == 0)
        arena[M.off + cell] = m;
      if ((m >> lane) & 1u)
      {
        const unsigned long long a = vbase + (unsigned long long)((long long)(x + lane) * su + (long long)y * sv);
        if (in[a] == 0)
          atomic_max_pixel(out + a, label);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
void
launch_dt_level(Launch & L, int ptype, const DtLevel & lv)
{
  if (lv.maxNodes <= 0)
    return;
  const int nodeGrid = std::max(1, std::min(lv.maxNodes, L.sms * 8));
  const int itemGrid = std::max(1, std::min(lv.itemCap, L.sms * 8));
  L.pre();
  k_dt_prep<<<nodeGrid, 256, 0, L.stream>>>(lv.nodes, lv.count, lv.work, lv.arena, lv.scratch, lv.scratchTop, lv.scratchCap,
                                             lv.items, lv.itemCount, lv.itemCap, lv.err);
  L.post("k_dt_prep");
  L.pre();
  switch (ptype)
  {
    case MCI_U8:
      k_dt_hist<uint8_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.items, lv.itemCount, lv.scratch, lv.hist, lv.histStride,
                                                          lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
      break;
    case MCI_U16:
      k_dt_hist<uint16_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.items, lv.itemCount, lv.scratch, lv.hist, lv.histStride,
                                                           lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
      break;
    default:
      k_dt_hist<int16_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.items, lv.itemCount, lv.scratch, lv.hist, lv.histStride,
                                                          lv.histSide, lv.bigHist, lv.bigCount, lv.nbig, lv.err);
  }
  L.post("k_dt_hist");
  L.pre();
  switch (ptype)
  {
    case MCI_U8:
      k_dt_scan<uint8_t><<<nodeGrid, 256, 0, L.stream>>>(lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
      break;
    case MCI_U16:
      k_dt_scan<uint16_t><<<nodeGrid, 256, 0, L.stream>>>(lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
      break;
    default:
      k_dt_scan<int16_t><<<nodeGrid, 256, 0, L.stream>>>(lv.work, lv.count, lv.hist, lv.histStride, lv.histSide, lv.bigHist);
  }
  L.post("k_dt_scan");
  L.pre();
  k_dt_median<<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.items, lv.itemCount, lv.scratch, (int)lv.dims[0], (int)lv.dims[1],
                                               (int)lv.dims[2], lv.err);
  L.post("k_dt_median");
  L.pre();
  k_dt_alloc<<<(lv.maxNodes + 127) / 128, 128, 0, L.stream>>>(lv.work, lv.count, lv.arenaTop, lv.arenaCap, lv.next, lv.nextCount,
                                                                lv.nextCap, lv.witems, lv.witemCount, lv.itemCap, lv.err);
  L.post("k_dt_alloc");
  L.pre();
  switch (ptype)
  {
    case MCI_U8:
      k_dt_emit<uint8_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                          (const uint8_t *)lv.in, (uint8_t *)lv.out, lv.dims[0], lv.dims[1],
                                                          lv.dims[2]);
      break;
    case MCI_U16:
      k_dt_emit<uint16_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                           (const uint16_t *)lv.in, (uint16_t *)lv.out, lv.dims[0], lv.dims[1],
                                                           lv.dims[2]);
      break;
    default:
      k_dt_emit<int16_t><<<itemGrid, 256, 0, L.stream>>>(lv.work, lv.witems, lv.witemCount, lv.scratch, lv.arena,
                                                          (const int16_t *)lv.in, (int16_t *)lv.out, lv.dims[0], lv.dims[1],
                                                          lv.dims[2]);
  }
  L.post("k_dt_emit");
}

} // namespace mci
