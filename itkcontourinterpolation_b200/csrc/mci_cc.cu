// mci_cc.cu -- batched 2-D connected components of the annotated slices, run-length based.
// (RegionedConnectedComponents, X:1224-1238; numbering of ITK's ConnectedComponentImageFilter: ids 1..n by the
//  raster position of each component's first pixel; the sums Centroid (X:1101-1134) needs.)
//
//   k_slice_bits   flat grid : extract the slice from the volume, == label, 32 pixels per ballot word
//   k_cc_runs      one CTA per slice, shared memory: row runs -> union-find over RUNS (the smaller run index is the
//                  parent, so a root is its component's raster-first run) -> consecutive ids in root order ->
//                  per-component count / sum u / sum v / bounding box
//   k_cc_paint     flat grid : component image (uint16 ids) from the run -> component table
//
// Runs, not pixels, are the union-find elements: a blob of 40 000 pixels is ~400 runs, the pointer chasing stays
// in shared memory, and its depth is bounded by the number of rows instead of by DRAM-latency-long pixel chains.
#include <algorithm>
#include "mci_kernels.cuh"
#include "mci_launch.h"

namespace mci
{

#define CC_THREADS 512
#define CC_SMEM_RUNS 6144 // runs of one slice kept in shared memory (x-extent + parent); beyond: global scratch

template <typename T>
__global__ void __launch_bounds__(256)
k_slice_bits(const T * __restrict__ in, const SliceDev * __restrict__ slices, const RowTile * __restrict__ tiles, uint32_t * bitsAll)
{
  pdl_wait();
  const RowTile tile = tiles[blockIdx.x];
  const SliceDev S = slices[tile.item];
  const int cw = (S.w + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int cell = warp; cell < tile.nrows * cw; cell += nwarps)
  {
    const int y = tile.row0 + cell / cw;
    const int cx = cell % cw;
    const int x = cx * 32 + lane;
    bool fg = false;
    if (x < S.w)
      fg = in[S.base + (u64)((long long)x * S.su + (long long)y * S.sv)] == (T)S.label;
    const uint32_t bits = __ballot_sync(0xFFFFFFFFu, fg);
    if (lane == 0)
      bitsAll[S.bitOff + (u64)y * cw + cx] = bits;
  }
}

__device__ __forceinline__ uint32_t
run_find(volatile uint32_t * parent, uint32_t r)
{
  uint32_t p;
  while ((p = parent[r]) != r)
    r = p;
  return r;
}

__device__ __forceinline__ void
run_union(uint32_t * parent, uint32_t a, uint32_t b)
{
  for (;;)
  {
    a = run_find(parent, a);
    b = run_find(parent, b);
    if (a == b)
      return;
    if (a < b)
    {
      uint32_t t = a;
      a = b;
      b = t;
    }
    uint32_t old = atomicMin(&parent[a], b);
    if (old == a)
      return;
    a = old;
  }
}

__global__ void __launch_bounds__(CC_THREADS)
k_cc_runs(const SliceDev * __restrict__ slices, const uint32_t * __restrict__ bitsAll, uint32_t * wordRunAll, uint32_t * runCompAll,
          uint32_t * runScratch, int fully, int * ncompArr, int * compBaseArr, int * totalComp, CompStat * stats, int statCap,
          int maxPerSlice, int * err)
{
  pdl_wait();
  extern __shared__ uint32_t smem[];
  __shared__ int s_wsum[CC_THREADS / 32];
  __shared__ int s_carry, s_total, s_base;
  const SliceDev S = slices[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = CC_THREADS / 32;
  const int cw = (S.w + 31) >> 5;
  const uint32_t * bits = bitsAll + S.bitOff;
  uint32_t * wordRun = wordRunAll + S.bitOff; // per word: index of the first run starting in it (global run numbering of the slice)
  uint32_t * runComp = runCompAll + S.runOff;
  // shared: rowBase[h + 1]
  int * rowBase = (int *)smem;

  // a group of GW lanes per row (GW = words per row rounded up to a power of two), 32 / GW rows per warp at once:
  // label bounding boxes are a few words wide, a whole warp per row would leave most lanes idle
  const int GW = cw <= 1 ? 1 : (cw <= 2 ? 2 : (cw <= 4 ? 4 : (cw <= 8 ? 8 : (cw <= 16 ? 16 : 32))));
  const int RPW = 32 / GW, sub = lane / GW, sl = lane - sub * GW;
  // ---- 1. run starts per row
  for (int yb = warp * RPW; yb < S.h; yb += NW * RPW)
  {
    const int y = yb + sub;
    const bool rowOk = y < S.h;
    int cnt = 0;
    uint32_t prevTop = 0;
    for (int wb = 0; wb < cw; wb += GW)
    {
      const int wi = wb + sl;
      const uint32_t w = (rowOk && wi < cw) ? bits[(size_t)y * cw + wi] : 0u;
      uint32_t below = __shfl_up_sync(0xFFFFFFFFu, w, 1, GW);
      if (sl == 0)
        below = prevTop;
      cnt += __popc(w & ~((w << 1) | (below >> 31)));
      prevTop = __shfl_sync(0xFFFFFFFFu, w, GW - 1, GW);
    }
    for (int o = GW >> 1; o > 0; o >>= 1)
      cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (sl == 0 && rowOk)
      rowBase[y] = cnt;
  }
  if (tid == 0)
    s_carry = 0;
  __syncthreads();
  // ---- 2. exclusive scan over rows
  for (int base = 0; base < S.h; base += CC_THREADS)
  {
    const int y = base + tid;
    const int v = y < S.h ? rowBase[y] : 0;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      int t = __shfl_up_sync(0xFFFFFFFFu, s, o);
      if (lane >= o)
        s += t;
    }
    if (lane == 31)
      s_wsum[warp] = s;
    __syncthreads();
    int off = s_carry;
    for (int k = 0; k < warp; ++k)
      off += s_wsum[k];
    if (y < S.h)
      rowBase[y] = off + s - v;
    __syncthreads();
    if (tid == CC_THREADS - 1)
      s_carry = off + s;
    __syncthreads();
  }
  const int R = s_carry;
  if (tid == 0)
    rowBase[S.h] = R;
  // run tables: x0 | x1 << 16, row, parent -- shared memory when they fit
  uint32_t * runX = (uint32_t *)(rowBase + S.h + 1);
  uint32_t * runRow = runX + CC_SMEM_RUNS;
  uint32_t * parent = runRow + CC_SMEM_RUNS;
  if (R > CC_SMEM_RUNS)
  {
    runX = runScratch + S.runOff * 3;
    runRow = runX + R;
    parent = runRow + R;
  }
  if (R > S.runCap)
  {
    if (tid == 0)
      atomicMax(err, ERR_COMPONENTS);
    if (tid == 0)
    {
      ncompArr[blockIdx.x] = 0;
      compBaseArr[blockIdx.x] = 0;
    }
    return;
  }
  __syncthreads();
  // ---- 3. run records (k-th start of a row pairs with its k-th end) and per-word run numbering
  for (int yb = warp * RPW; yb < S.h; yb += NW * RPW)
  {
    const int y = yb + sub;
    const bool rowOk = y < S.h;
    int nS = rowOk ? rowBase[y] : 0, nE = nS;
    uint32_t prevTop = 0;
    for (int wb = 0; wb < cw; wb += GW)
    {
      const int wi = wb + sl;
      const bool wordOk = rowOk && wi < cw;
      const uint32_t w = wordOk ? bits[(size_t)y * cw + wi] : 0u;
      uint32_t below = __shfl_up_sync(0xFFFFFFFFu, w, 1, GW);
      if (sl == 0)
        below = prevTop;
      uint32_t above = __shfl_down_sync(0xFFFFFFFFu, w, 1, GW);
      if (sl == GW - 1)
        above = (rowOk && wb + GW < cw) ? bits[(size_t)y * cw + wb + GW] : 0u;
      const uint32_t starts = w & ~((w << 1) | (below >> 31));
      const uint32_t ends = w & ~((w >> 1) | (above << 31));
      int cs = __popc(starts), ce = __popc(ends);
      int is = cs, ie = ce;
      for (int o = 1; o < GW; o <<= 1)
      {
        int a = __shfl_up_sync(0xFFFFFFFFu, is, o, GW), b = __shfl_up_sync(0xFFFFFFFFu, ie, o, GW);
        if (sl >= o)
        {
          is += a;
          ie += b;
        }
      }
      int ks = nS + is - cs, ke = nE + ie - ce;
      if (wordOk)
        wordRun[(size_t)y * cw + wi] = (uint32_t)ks;
      uint32_t m = starts;
      while (m)
      {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        runX[ks] = (uint32_t)(wi * 32 + b); // x1 is OR-ed in below
        runRow[ks] = (uint32_t)y;
        parent[ks] = (uint32_t)ks;
        ++ks;
      }
      nS += __shfl_sync(0xFFFFFFFFu, is, GW - 1, GW);
      prevTop = __shfl_sync(0xFFFFFFFFu, w, GW - 1, GW);
      // ends are written after all starts of the row are in place (same warp, program order)
      __syncwarp();
      m = ends;
      while (m)
      {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        atomicOr(&runX[ke], (uint32_t)(wi * 32 + b) << 16);
        ++ke;
      }
      nE += __shfl_sync(0xFFFFFFFFu, ie, GW - 1, GW);
    }
  }
  __syncthreads();
  // ---- 4. link every run to the runs of the previous row it touches (face- or fully connected)
  const int reach = fully ? 1 : 0;
  for (int r = tid; r < R; r += CC_THREADS)
  {
    const int y = (int)runRow[r];
    if (y == 0)
      continue;
    const int x0 = (int)(runX[r] & 0xFFFFu) - reach, x1 = (int)(runX[r] >> 16) + reach;
    for (int q = rowBase[y - 1]; q < rowBase[y]; ++q)
    {
      const int X0 = (int)(runX[q] & 0xFFFFu), X1 = (int)(runX[q] >> 16);
      if (X0 > x1)
        break;
      if (X1 >= x0)
        run_union(parent, (uint32_t)r, (uint32_t)q);
    }
  }
  __syncthreads();
  // ---- 5. roots -> consecutive ids in run order (= raster order of each component's first pixel)
  if (tid == 0)
    s_carry = 0;
  __syncthreads();
  for (int base = 0; base < R; base += CC_THREADS)
  {
    const int r = base + tid;
    const int isRoot = (r < R && parent[r] == (uint32_t)r) ? 1 : 0;
    int s = isRoot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      int t = __shfl_up_sync(0xFFFFFFFFu, s, o);
      if (lane >= o)
        s += t;
    }
    if (lane == 31)
      s_wsum[warp] = s;
    __syncthreads();
    int off = s_carry;
    for (int k = 0; k < warp; ++k)
      off += s_wsum[k];
    if (isRoot)
      runComp[r] = (uint32_t)(off + s); // 1-based id
    __syncthreads();
    if (tid == CC_THREADS - 1)
      s_carry = off + s;
    __syncthreads();
  }
  const int ncomp = s_carry;
  if (tid == 0)
  {
    int base = atomicAdd(totalComp, ncomp);
    if (ncomp > maxPerSlice)
      atomicMax(err, ERR_PIXELTYPE);
    if (base + ncomp > statCap)
    {
      atomicMax(err, ERR_COMPONENTS);
      base = -1;
    }
    s_base = base;
    ncompArr[blockIdx.x] = ncomp;
    compBaseArr[blockIdx.x] = base < 0 ? 0 : base;
  }
  __syncthreads();
  const int cbase = s_base;
  for (int r = tid; r < R; r += CC_THREADS)
    if (parent[r] != (uint32_t)r)
      runComp[r] = runComp[run_find(parent, (uint32_t)r)];
  if (cbase < 0)
    return;
  // ---- 6. per-component statistics from the runs
  for (int k = tid; k < ncomp; k += CC_THREADS)
  {
    CompStat s;
    s.count = 0;
    s.sumU = 0;
    s.sumV = 0;
    s.minU = 0x7FFFFFFF;
    s.minV = 0x7FFFFFFF;
    s.maxU = -0x7FFFFFFF;
    s.maxV = -0x7FFFFFFF;
    stats[cbase + k] = s;
  }
  __syncthreads();
  for (int r = tid; r < R; r += CC_THREADS)
  {
    const int x0 = (int)(runX[r] & 0xFFFFu), x1 = (int)(runX[r] >> 16), y = (int)runRow[r];
    const long long len = x1 - x0 + 1;
    const long long u = S.u0 + x0, v = S.v0 + y;
    CompStat * st = stats + cbase + (int)runComp[r] - 1;
    atomicAdd((unsigned long long *)&st->count, (unsigned long long)len);
    atomicAdd((unsigned long long *)&st->sumU, (unsigned long long)(u * len + len * (len - 1) / 2));
    atomicAdd((unsigned long long *)&st->sumV, (unsigned long long)(v * len));
    atomicMin(&st->minU, (int)u);
    atomicMax(&st->maxU, (int)(u + len - 1));
    atomicMin(&st->minV, (int)v);
    atomicMax(&st->maxV, (int)v);
  }
}

__global__ void __launch_bounds__(256)
k_cc_paint(const SliceDev * __restrict__ slices, const RowTile * __restrict__ tiles, const uint32_t * __restrict__ bitsAll,
           const uint32_t * __restrict__ wordRunAll, const uint32_t * __restrict__ runCompAll, uint16_t * connAll)
{
  pdl_wait();
  const RowTile tile = tiles[blockIdx.x];
  const SliceDev S = slices[tile.item];
  const int cw = (S.w + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int cell = warp; cell < tile.nrows * cw; cell += nwarps)
  {
    const int y = tile.row0 + cell / cw;
    const int cx = cell % cw;
    const int x = cx * 32 + lane;
    const size_t widx = S.bitOff + (size_t)y * cw + cx;
    const uint32_t w = __ldg(bitsAll + widx);
    uint32_t id = 0;
    if (w)
    {
      const uint32_t below = cx > 0 ? __ldg(bitsAll + widx - 1) : 0u;
      const uint32_t starts = w & ~((w << 1) | (below >> 31));
      if ((w >> lane) & 1u)
      {
        // the run holding this pixel: the last one that started at or before it in the row
        const uint32_t r = __ldg(wordRunAll + widx) + __popc(starts & (0xFFFFFFFFu >> (31 - lane))) - 1u;
        id = __ldg(runCompAll + S.runOff + r);
      }
    }
    if (x < S.w)
      connAll[S.pixOff + (u64)y * S.w + x] = (uint16_t)id;
  }
}

void
launch_cc(Launch & L, int ptype, const void * in, const SliceDev * slices, int nslices, const RowTile * tiles, int ntiles,
          uint32_t * bits, uint32_t * wordRun, uint32_t * runComp, uint32_t * runScratch, uint16_t * conn, int * ncomp,
          int * compBase, int * total, CompStat * stats, int statCap, int fully, int maxRows, int * err)
{
  if (nslices == 0 || ntiles == 0)
    return;
  L.pre("k_slice_bits");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_slice_bits<uint8_t>, ntiles, 256, 0, L.stream, (const uint8_t *)in, slices, tiles, bits);
      break;
    case MCI_U16:
      launch_pdl(k_slice_bits<uint16_t>, ntiles, 256, 0, L.stream, (const uint16_t *)in, slices, tiles, bits);
      break;
    default:
      launch_pdl(k_slice_bits<int16_t>, ntiles, 256, 0, L.stream, (const int16_t *)in, slices, tiles, bits);
  }
  L.post("k_slice_bits");
  L.ensure_max_smem(k_cc_runs, 0, L.maxSmem - 2048);
  const int smemBytes = ((maxRows + 1) + 3 * CC_SMEM_RUNS) * 4;
  const int maxPer = ptype == MCI_U8 ? 255 : (ptype == MCI_U16 ? 65535 : 32767);
  L.pre("k_cc_runs");
  launch_pdl(k_cc_runs, nslices, CC_THREADS, smemBytes, L.stream, slices, bits, wordRun, runComp, runScratch, fully, ncomp, compBase, total,
                                                           stats, statCap, maxPer, err);
  L.post("k_cc_runs");
  L.pre("k_cc_paint");
  launch_pdl(k_cc_paint, ntiles, 256, 0, L.stream, slices, tiles, bits, wordRun, runComp, conn);
  L.post("k_cc_paint");
}

} // namespace mci
