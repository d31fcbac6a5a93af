// mci_dil.cu -- one Interpolate1to1 call per CTA with the DILATION median (UseDistanceTransform = off):
// translate + binarise + AND (X:570-666), FindMedianImageDilations (X:340-388) on bit-packed shapes held in shared
// memory, clip + write with the max rule (X:679-764), spawn the two children (X:766-792).
//
// No sequence image is ever stored.  |seq[x] symdiff iMask| depends only on how many pixels of i\j and of j\i
// joined the constrained dilation at each step (intersection pixels are in every image), so the two sequences run
// once to fill two step histograms -- BOTH AT ONCE, one barrier per step -- and, once the best x is known, run again
// up to the two chosen steps to rebuild seq[x] = D^i_si | D^j_sj.  A word of a sequence can only change when it or
// one of its eight neighbouring words changed in the previous step; changed words wake their neighbourhood in a
// bitmap and every other word is skipped.
#include <algorithm>
#include <type_traits>
#include "mci_kernels.cuh"
#include "mci_launch.h"

namespace mci
{

#define DIL_THREADS 512
#define DIL_STEPS 65536 // capacity of the per-CTA step histograms

struct DilShared
{
  int anyX;
  int bx0, by0, bx1, by1;
  unsigned long long midOff;
  long long redA[DIL_THREADS / 32], redB[DIL_THREADS / 32], bestD[DIL_THREADS / 32];
  int bestI[DIL_THREADS / 32];
  int result;
  unsigned flag[3]; // rotating "something changed" words: bit 0 = i sequence, bit 1 = j sequence
  unsigned cntI[3], cntJ[3]; // rotating per-step counts of pixels that joined outside the other shape
};

// wakes word idx and its two row neighbours (a superset of what can be reached is always safe)
__device__ __forceinline__ void
dil_wake_row3(uint32_t * wake, int idx, int words)
{
  const int lo = max(idx - 1, 0), hi = min(idx + 1, words - 1);
  const int w0 = lo >> 5, w1 = hi >> 5;
  if (w0 == w1)
    atomicOr(&wake[w0], ((0xFFFFFFFFu >> (31 - (hi & 31))) & (0xFFFFFFFFu << (lo & 31))));
  else
  {
    atomicOr(&wake[w0], 0xFFFFFFFFu << (lo & 31));
    atomicOr(&wake[w1], 0xFFFFFFFFu >> (31 - (hi & 31)));
  }
}

// One sequence's share of one step for word idx: returns true if the word changed.
__device__ __forceinline__ bool
dil_step_word(const uint32_t * cur, uint32_t * nxt, const uint32_t * mask, const uint32_t * other, uint32_t * wakeNxt, int idx, int h,
              int pitch, int words, bool ball, unsigned * histSlot, unsigned short * arrival, unsigned short step)
{
  const int r = idx / pitch, wi = idx - r * pitch;
  const uint32_t old = cur[idx];
  const uint32_t m = mask[idx];
  uint32_t v = old;
  if (m & ~old)
    v = old | (dilate_word(cur, h, pitch, r, wi, ball) & m);
  nxt[idx] = v;
  const uint32_t diff = v & ~old;
  if (!diff)
    return false;
  if (histSlot)
    *histSlot += (unsigned)__popc(diff & ~other[idx]); // per-thread counter, reduced once per step
  if (arrival)
  {
    // the step at which each pixel joins: lets the chosen image be rebuilt by thresholding instead of re-running
    uint32_t m2 = diff;
    while (m2)
    {
      const int bpos = __ffs(m2) - 1;
      m2 &= m2 - 1;
      arrival[(size_t)idx * 32 + bpos] = step;
    }
  }
  if (r > 0)
    dil_wake_row3(wakeNxt, idx - pitch, words);
  dil_wake_row3(wakeNxt, idx, words);
  if (r + 1 < h)
    dil_wake_row3(wakeNxt, idx + pitch, words);
  return true;
}

// Both constrained dilation sequences (GenerateDilationSequence X:322-338, Dilate1 X:253-320) advanced together.
// stopI / stopJ < 0: run until nothing grows, recording in hI[s] / hJ[s] how many pixels outside the other shape
// joined at step s; otherwise stop sequence i (j) after stopI (stopJ) growth steps.  On return Di[ci] / Dj[cj]
// hold the last images; K = number of images of each sequence.
__device__ void
dil_run_pair(DilShared & sh, const uint32_t * Ib, const uint32_t * Jb, uint32_t * Di0, uint32_t * Di1, uint32_t * Dj0, uint32_t * Dj1,
             uint32_t * wake /* 6 bitmaps */, int h, int pitch, bool ball, int stopI, int stopJ, unsigned * hI, unsigned * hJ,
             unsigned short * arrI, unsigned short * arrJ, int & Ki, int & Kj, int & ci, int & cj)
{
  const int words = h * pitch, nbw = (words + 31) >> 5;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < words; idx += DIL_THREADS)
  {
    const uint32_t x = Ib[idx] & Jb[idx];
    Di0[idx] = x;
    Di1[idx] = x;
    Dj0[idx] = x;
    Dj1[idx] = x;
  }
  // wake bitmaps: [seq][3 rotating][nbw]; step 1 reads slot 0 (everything awake), writes slot 1, clears slot 2
  for (int k = tid; k < nbw; k += DIL_THREADS)
  {
    wake[0 * nbw + k] = 0xFFFFFFFFu;
    wake[1 * nbw + k] = 0u;
    wake[2 * nbw + k] = 0u;
    wake[3 * nbw + k] = 0xFFFFFFFFu;
    wake[4 * nbw + k] = 0u;
    wake[5 * nbw + k] = 0u;
  }
  if (tid == 0)
  {
    sh.flag[0] = sh.flag[1] = sh.flag[2] = 0u;
    sh.cntI[0] = sh.cntI[1] = sh.cntI[2] = 0u;
    sh.cntJ[0] = sh.cntJ[1] = sh.cntJ[2] = 0u;
  }
  __syncthreads();
  int stepsI = 0, stepsJ = 0;
  bool runI = stopI != 0, runJ = stopJ != 0;
  ci = 0;
  cj = 0;
  for (int s = 0; runI || runJ; ++s)
  {
    const int slot = s % 3, slotN = (s + 1) % 3, slotC = (s + 2) % 3;
    unsigned mine = 0;
    uint32_t * curI = ci ? Di1 : Di0;
    uint32_t * nxtI = ci ? Di0 : Di1;
    uint32_t * curJ = cj ? Dj1 : Dj0;
    uint32_t * nxtJ = cj ? Dj0 : Dj1;
    unsigned myCntI = 0, myCntJ = 0;
    unsigned * slotHI = hI ? &myCntI : nullptr;
    unsigned * slotHJ = hJ ? &myCntJ : nullptr;
    // which of my words are awake: independent loads first (one bitmap word serves a whole warp), then only
    // the awake ones are visited -- dormant words cost no dependent shared-memory round trip
    unsigned awI = 0, awJ = 0;
    {
      int j = 0;
      for (int idx = tid; idx < words && j < 32; idx += DIL_THREADS, ++j)
      {
        if (runI)
          awI |= ((wake[slot * nbw + (idx >> 5)] >> (idx & 31)) & 1u) << j;
        if (runJ)
          awJ |= ((wake[(3 + slot) * nbw + (idx >> 5)] >> (idx & 31)) & 1u) << j;
      }
    }
    while (awI)
    {
      const int j = __ffs(awI) - 1;
      awI &= awI - 1;
      if (dil_step_word(curI, nxtI, Ib, Jb, wake + slotN * nbw, tid + j * DIL_THREADS, h, pitch, words, ball, slotHI, arrI,
                        (unsigned short)(stepsI + 1)))
        mine |= 1u;
    }
    while (awJ)
    {
      const int j = __ffs(awJ) - 1;
      awJ &= awJ - 1;
      if (dil_step_word(curJ, nxtJ, Jb, Ib, wake + (3 + slotN) * nbw, tid + j * DIL_THREADS, h, pitch, words, ball, slotHJ, arrJ,
                        (unsigned short)(stepsJ + 1)))
        mine |= 2u;
    }
    for (int idx = tid + 32 * DIL_THREADS; idx < words; idx += DIL_THREADS) // shapes beyond 32 words per thread
    {
      if (runI && ((wake[slot * nbw + (idx >> 5)] >> (idx & 31)) & 1u) &&
          dil_step_word(curI, nxtI, Ib, Jb, wake + slotN * nbw, idx, h, pitch, words, ball, slotHI, arrI, (unsigned short)(stepsI + 1)))
        mine |= 1u;
      if (runJ && ((wake[(3 + slot) * nbw + (idx >> 5)] >> (idx & 31)) & 1u) &&
          dil_step_word(curJ, nxtJ, Jb, Ib, wake + (3 + slotN) * nbw, idx, h, pitch, words, ball, slotHJ, arrJ, (unsigned short)(stepsJ + 1)))
        mine |= 2u;
    }
    for (int k = tid; k < nbw; k += DIL_THREADS)
    {
      wake[slotC * nbw + k] = 0u; // read two steps ago, written from the next step on
      wake[(3 + slotC) * nbw + k] = 0u;
    }
    if (mine)
      atomicOr(&sh.flag[slot], mine);
    if (myCntI)
      atomicAdd(&sh.cntI[slot], myCntI);
    if (myCntJ)
      atomicAdd(&sh.cntJ[slot], myCntJ);
    if (tid == 0)
    {
      sh.flag[slotN] = 0u;
      sh.cntI[slotN] = 0u;
      sh.cntJ[slotN] = 0u;
    }
    __syncthreads();
    const unsigned f = sh.flag[slot];
    if (tid == 0)
    {
      if (hI && runI && (f & 1u) && stepsI + 1 < DIL_STEPS)
        hI[stepsI + 1] = sh.cntI[slot];
      if (hJ && runJ && (f & 2u) && stepsJ + 1 < DIL_STEPS)
        hJ[stepsJ + 1] = sh.cntJ[slot];
    }
    if (runI)
    {
      if (f & 1u)
      {
        ++stepsI;
        ci ^= 1;
        if (stopI > 0 && stepsI >= stopI)
          runI = false;
      }
      else
        runI = false; // nothing grew: the last image was a duplicate (X:334-336)
    }
    if (runJ)
    {
      if (f & 2u)
      {
        ++stepsJ;
        cj ^= 1;
        if (stopJ > 0 && stepsJ >= stopJ)
          runJ = false;
      }
      else
        runJ = false;
    }
  }
  Ki = stepsI > 0 ? stepsI : 1; // [D1] alone when nothing grows (X:328-336)
  Kj = stepsJ > 0 ? stepsJ : 1;
  __syncthreads();
}

__device__ __forceinline__ void
dil_block_scan2(DilShared & sh, long long a, long long b, long long & pa, long long & pb, long long & totA, long long & totB)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long sa = a, sb = b;
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
  __syncthreads();
  if (lane == 31)
  {
    sh.redA[warp] = sa;
    sh.redB[warp] = sb;
  }
  __syncthreads();
  long long oa = 0, ob = 0;
  totA = 0;
  totB = 0;
  for (int k = 0; k < DIL_THREADS / 32; ++k)
  {
    if (k < warp)
    {
      oa += sh.redA[k];
      ob += sh.redB[k];
    }
    totA += sh.redA[k];
    totB += sh.redB[k];
  }
  pa = oa + sa - a;
  pb = ob + sb - b;
}

__device__ __forceinline__ int
dil_block_argmin(DilShared & sh, long long diff, int idx)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    long long od = __shfl_xor_sync(0xFFFFFFFFu, diff, o);
    int oi = __shfl_xor_sync(0xFFFFFFFFu, idx, o);
    if (od < diff || (od == diff && oi < idx))
    {
      diff = od;
      idx = oi;
    }
  }
  __syncthreads();
  if (lane == 0)
  {
    sh.bestD[warp] = diff;
    sh.bestI[warp] = idx;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    long long d = sh.bestD[0];
    int i = sh.bestI[0];
    for (int k = 1; k < DIL_THREADS / 32; ++k)
      if (sh.bestD[k] < d || (sh.bestD[k] == d && sh.bestI[k] < i))
      {
        d = sh.bestD[k];
        i = sh.bestI[k];
      }
    sh.result = i;
  }
  __syncthreads();
  return sh.result;
}

template <typename T>
__global__ void __launch_bounds__(DIL_THREADS, 2)
k_dil_nodes(const Node * __restrict__ nodes, const int * __restrict__ count, Node * next, int * nextCount, int nextCap, uint32_t * arena,
            unsigned long long * arenaTop, unsigned long long arenaCap, const T * __restrict__ in, T * out, long long VX,
            long long VY, long long VZ, int ball, unsigned * stepHistAll, uint32_t * slabAll, unsigned long long slabWords,
            unsigned short * arrAll, unsigned long long arrPixels, int smemWords, EmitRec * emit, int * emitCount, int emitCap, unsigned long long emitBase, int * err)
{
  pdl_wait();
  extern __shared__ uint32_t smem[];
  __shared__ DilShared sh;
  unsigned * hI = stepHistAll + (size_t)blockIdx.x * 2 * DIL_STEPS; // hI[s], s = 1..Ki: pixels of i\\j joining at step s
  unsigned * hJ = hI + DIL_STEPS;
  const int tid = threadIdx.x;
  const int nNodes = *count;
  for (int nidx = blockIdx.x; nidx < nNodes; nidx += gridDim.x)
  {
    const Node nd = nodes[nidx];
    // ---- translation halving with carry (X:570-608)
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
    // ---- working region: union of the two translated shapes' regions
    const int ux0 = min(nd.A.x0 + iT[0], nd.B.x0 + jT[0]), uy0 = min(nd.A.y0 + iT[1], nd.B.y0 + jT[1]);
    const int ux1 = max(nd.A.x0 + nd.A.w + iT[0], nd.B.x0 + nd.B.w + jT[0]);
    const int uy1 = max(nd.A.y0 + nd.A.h + iT[1], nd.B.y0 + nd.B.h + jT[1]);
    const int w = ux1 - ux0, h = uy1 - uy0;
    const int pitch = (w + 31) >> 5;
    const int words = h * pitch, nbw = (words + 31) >> 5;
    const long long need = 6ll * words + 6ll * nbw;
    uint32_t * base;
    if (need <= smemWords)
      base = smem;
    else if ((unsigned long long)need <= slabWords)
      base = slabAll + (size_t)blockIdx.x * slabWords;
    else
    {
      if (tid == 0)
        atomicMax(err, ERR_SCRATCH);
      continue;
    }
    uint32_t * Ib = base;
    uint32_t * Jb = base + words;
    uint32_t * Di0 = base + 2 * (size_t)words;
    uint32_t * Di1 = base + 3 * (size_t)words;
    uint32_t * Dj0 = base + 4 * (size_t)words;
    uint32_t * Dj1 = base + 5 * (size_t)words;
    uint32_t * wake = base + 6 * (size_t)words;
    // arrival-step maps (global, one pair per CTA); shapes too large for them fall back to re-running the sequences
    const bool haveArr = arrAll != nullptr && (unsigned long long)words * 32 <= arrPixels;
    unsigned short * arrI = haveArr ? arrAll + (size_t)blockIdx.x * 2 * arrPixels : nullptr;
    unsigned short * arrJ = haveArr ? arrI + arrPixels : nullptr;
    __syncthreads(); // previous node done with shared state
    if (tid == 0)
    {
      sh.anyX = 0;
      sh.bx0 = sh.by0 = 0x7FFFFFFF;
      sh.bx1 = sh.by1 = -0x7FFFFFFF;
    }
    __syncthreads();
    // ---- translate, binarise, intersect (X:610-666)
    {
      int any = 0;
      for (int idx = tid; idx < words; idx += DIL_THREADS)
      {
        const int r = idx / pitch, wi = idx - r * pitch;
        const int y = uy0 + r, x = ux0 + 32 * wi;
        const uint32_t a = mask_word(arena, nd.A, y - iT[1], x - iT[0]);
        const uint32_t b = mask_word(arena, nd.B, y - jT[1], x - jT[0]);
        Ib[idx] = a;
        Jb[idx] = b;
        any |= (a & b) != 0u;
      }
      if (any)
        sh.anyX = 1;
    }
    __syncthreads();
    if (!sh.anyX)
      continue; // empty intersection: the median is empty, nothing below can be written
    // ---- FindMedianImageDilations (X:340-388)
    int Ki, Kj, ci, cj;
    dil_run_pair(sh, Ib, Jb, Di0, Di1, Dj0, Dj1, wake, h, pitch, ball != 0, -1, -1, hI, hJ, arrI, arrJ, Ki, Kj, ci, cj);
    if (Ki >= DIL_STEPS || Kj >= DIL_STEPS)
    {
      if (tid == 0)
        atomicMax(err, ERR_SCRATCH);
      Ki = min(Ki, DIL_STEPS - 1);
      Kj = min(Kj, DIL_STEPS - 1);
    }
    // totals of i\\j and j\\i (pixels a sequence never reaches still count in the symmetric differences)
    long long ni = 0, nj = 0;
    for (int idx = tid; idx < words; idx += DIL_THREADS)
    {
      ni += __popc(Ib[idx] & ~Jb[idx]);
      nj += __popc(Jb[idx] & ~Ib[idx]);
    }
    long long d0, d1, Ni, Nj;
    dil_block_scan2(sh, ni, nj, d0, d1, Ni, Nj);
    // inclusive prefix sums of the step histograms, in place
    const int n = max(Ki, Kj) + 1;
    {
      const int chunk = (n + DIL_THREADS - 1) / DIL_THREADS;
      const int lo = min(n, tid * chunk), hiEnd = min(n, lo + chunk);
      long long cI = 0, cJ = 0;
      for (int k = lo; k < hiEnd; ++k)
      {
        cI += (k >= 1 && k <= Ki) ? hI[k] : 0u;
        cJ += (k >= 1 && k <= Kj) ? hJ[k] : 0u;
      }
      long long pi, pj, ti, tj;
      dil_block_scan2(sh, cI, cJ, pi, pj, ti, tj);
      long long si = pi, sj = pj;
      for (int k = lo; k < hiEnd; ++k)
      {
        si += (k >= 1 && k <= Ki) ? hI[k] : 0u;
        sj += (k >= 1 && k <= Kj) ? hJ[k] : 0u;
        hI[k] = (unsigned)si;
        hJ[k] = (unsigned)sj;
      }
    }
    __syncthreads();
    // seq[x] (X:349-371): the i-sequence is reversed BEFORE the swap that makes iSeq the longer one
    const bool swapped = Ki < Kj;
    const int L = swapped ? Kj : Ki;
    const float ratio = __fdiv_rn((float)(swapped ? Ki : Kj), (float)L);
    long long bestDiff = 0x7FFFFFFFFFFFFFFFll;
    int bestIdx = 0x7FFFFFFF;
    for (int x = tid; x < L; x += DIL_THREADS)
    {
      const unsigned xj = __float2uint_rz(__fmul_rn(ratio, (float)x));
      const int si = swapped ? Ki - (int)xj : Ki - x;
      const int sj = swapped ? x + 1 : (int)xj + 1;
      const long long cI = hI[min(si, Ki)], cJ = hJ[min(sj, Kj)];
      const long long iS = cJ + (Ni - cI); // |seq symdiff iMask|
      const long long jS = cI + (Nj - cJ); // |seq symdiff jMask|
      const long long d = llabs(iS - jS);
      if (d < bestDiff)
      {
        bestDiff = d;
        bestIdx = x;
      }
    }
    const int bx = dil_block_argmin(sh, bestDiff, bestIdx);
    // (the reference starts from min = number of region pixels with a strict "<": any x beats it)
    const unsigned bxj = __float2uint_rz(__fmul_rn(ratio, (float)bx));
    const int si = swapped ? Ki - (int)bxj : Ki - bx;
    const int sj = swapped ? bx + 1 : (int)bxj + 1;
    for (int k = tid; k < n; k += DIL_THREADS)
    {
      hI[k] = 0u; // leave the tables clean for the next node
      hJ[k] = 0u;
    }
    __syncthreads();
    // ---- rebuild the chosen image: D^i_si | D^j_sj (images are numbered from 1; image s = s growth steps,
    // image 1 of a sequence that never grows is the intersection itself)
    uint32_t * Sb;
    if (haveArr)
    {
      // threshold the arrival steps: pixel p of the final i image belongs to D^i_si iff it is an intersection pixel
      // or joined at a step <= si (same for j)
      const uint32_t * fi = ci ? Di1 : Di0;
      const uint32_t * fj = cj ? Dj1 : Dj0;
      Sb = ci ? Di0 : Di1;
      for (int idx = tid; idx < words; idx += DIL_THREADS)
      {
        const uint32_t x = Ib[idx] & Jb[idx];
        uint32_t m = x;
        uint32_t ri = fi[idx] & ~x, rj = fj[idx] & ~x;
        while (ri)
        {
          const int bpos = __ffs(ri) - 1;
          ri &= ri - 1;
          if ((int)arrI[(size_t)idx * 32 + bpos] <= si)
            m |= 1u << bpos;
        }
        while (rj)
        {
          const int bpos = __ffs(rj) - 1;
          rj &= rj - 1;
          if ((int)arrJ[(size_t)idx * 32 + bpos] <= sj)
            m |= 1u << bpos;
        }
        Sb[idx] = m;
      }
    }
    else
    {
      int k1, k2;
      dil_run_pair(sh, Ib, Jb, Di0, Di1, Dj0, Dj1, wake, h, pitch, ball != 0, si, sj, nullptr, nullptr, nullptr, nullptr, k1, k2, ci, cj);
      Sb = ci ? Di0 : Di1; // the i buffer that is NOT current receives the median
      const uint32_t * fi = ci ? Di1 : Di0;
      const uint32_t * fj = cj ? Dj1 : Dj0;
      for (int idx = tid; idx < words; idx += DIL_THREADS)
        Sb[idx] = fi[idx] | fj[idx];
    }
    __syncthreads();
    // ---- clip to the slice (X:679-712) and crop to the median's bounding box
    const int d0a = nd.axis == 0 ? 1 : 0, d1a = nd.axis == 2 ? 1 : 2;
    const long long dimsV[3] = { VX, VY, VZ };
    const int SU = (int)dimsV[d0a], SV = (int)dimsV[d1a];
    {
      int bx0 = 0x7FFFFFFF, bx1 = -0x7FFFFFFF, by0 = 0x7FFFFFFF, by1 = -0x7FFFFFFF;
      for (int idx = tid; idx < words; idx += DIL_THREADS)
      {
        const int r = idx / pitch, wi = idx - r * pitch;
        const int y = uy0 + r;
        uint32_t m = Sb[idx];
        if (!m || y < 0 || y >= SV)
          continue;
        const int xa = ux0 + 32 * wi;
        if (xa < 0)
          m = (xa <= -32) ? 0u : (m & (0xFFFFFFFFu << (-xa)));
        if (xa + 32 > SU)
          m = (xa >= SU) ? 0u : (m & (0xFFFFFFFFu >> (xa + 32 - SU)));
        if (!m)
          continue;
        bx0 = min(bx0, xa + __ffs(m) - 1);
        bx1 = max(bx1, xa + 31 - __clz(m));
        by0 = min(by0, y);
        by1 = max(by1, y);
      }
      if (by1 >= by0)
      {
        atomicMin(&sh.bx0, bx0);
        atomicMax(&sh.bx1, bx1);
        atomicMin(&sh.by0, by0);
        atomicMax(&sh.by1, by1);
      }
    }
    __syncthreads();
    if (sh.by1 < sh.by0)
      continue; // nothing inside the image
    MaskRef M;
    M.x0 = sh.bx0;
    M.y0 = sh.by0;
    M.w = sh.bx1 - sh.bx0 + 1;
    M.h = sh.by1 - sh.by0 + 1;
    M.pitch = (M.w + 31) >> 5;
    M.pad = 0;
    const int mwords = M.h * M.pitch;
    if (tid == 0)
      sh.midOff = atomicAdd(arenaTop, (unsigned long long)mwords);
    __syncthreads();
    M.off = sh.midOff;
    if (M.off + (unsigned long long)mwords > arenaCap)
    {
      if (tid == 0)
        atomicMax(err, ERR_ARENA);
      continue;
    }
    if (emit && tid == 0)
    {
      const int e = atomicAdd(emitCount, 1);
      if (e < emitCap)
      {
        EmitRec r;
        r.M = M;
        r.M.off = M.off - emitBase;
        r.axis = nd.axis;
        r.label = nd.label;
        r.mid = mid;
        r.pad = 0;
        emit[e] = r;
        if (nd.slot0 >= 0)
          emit_slot_table(emit, emitCap)[nd.slot0 + mid - nd.lo - 1] = e;
      }
      else
        atomicMax(err, ERR_NODE_LIST);
    }
    // ---- store the mid shape (X:714-729); its volume write with the max rule (X:743-764) is done for all nodes
    // at once by k_apply_masks after the last level
    for (int cell = tid; cell < mwords; cell += DIL_THREADS)
    {
      const int r = cell / M.pitch, wi = cell - r * M.pitch;
      const int y = M.y0 + r, x = M.x0 + 32 * wi;
      const uint32_t * row = Sb + (size_t)(y - uy0) * pitch;
      const int rx = x - ux0; // >= 0
      const int sw = rx >> 5, shf = rx & 31;
      const uint32_t lo = row[sw];
      const uint32_t hi = sw + 1 < pitch ? row[sw + 1] : 0u;
      uint32_t m = __funnelshift_r(lo, hi, shf);
      const int valid = M.w - 32 * wi;
      if (valid < 32)
        m &= (1u << valid) - 1u;
      arena[M.off + cell] = m;
    }
    // ---- children (X:766-792)
    if (tid == 0 && abs(nd.i - nd.j) > 2)
    {
      const bool first = abs(nd.i - mid) > 1, second = abs(nd.j - mid) > 1;
      const int nnew = (first ? 1 : 0) + (second ? 1 : 0);
      if (nnew)
      {
        int pos = atomicAdd(nextCount, nnew);
        if (pos + nnew > nextCap)
          atomicMax(err, ERR_NODE_LIST);
        else
        {
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
            next[pos++] = c;
          }
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
            next[pos++] = c;
          }
        }
      }
    }
  }
}

template <typename T>
static void
launch_dil_t(Launch & L, const Node * nodes, const int * count, int maxNodes, Node * next, int * nextCount, int nextCap,
             uint32_t * arena, unsigned long long * arenaTop, unsigned long long arenaCap, const void * in, void * out,
             const long long dims[3], int ball, unsigned * stepHist, uint32_t * slab, unsigned long long slabWords,
             unsigned short * arr, unsigned long long arrPixels, int grid, int smemBytes, EmitRec * emit, int * emitCount, int emitCap, unsigned long long emitBase, int * err)
{
  L.ensure_max_smem(k_dil_nodes<T>, 3 + (int)sizeof(T) + (std::is_signed<T>::value ? 1 : 0), L.maxSmem - 2048);
  const int g = std::max(1, std::min(grid, maxNodes));
  L.pre("k_dil_nodes");
  launch_pdl(k_dil_nodes<T>, g, DIL_THREADS, smemBytes, L.stream, nodes, count, next, nextCount, nextCap, arena, arenaTop, arenaCap,
                                                          (const T *)in, (T *)out, dims[0], dims[1], dims[2], ball, stepHist, slab,
                                                          slabWords, arr, arrPixels, smemBytes / 4, emit, emitCount, emitCap, emitBase, err);
  L.post("k_dil_nodes");
}

void
launch_nodes(Launch & L, int ptype, const Node * nodes, const int * count, int maxNodes, Node * next, int * nextCount,
             int nextCap, uint32_t * arena, unsigned long long * arenaTop, unsigned long long arenaCap, const void * in,
             void * out, const long long dims[3], int useDT, int ball, unsigned * bigHist, uint32_t * slab,
             unsigned long long slabWords, unsigned short * arr, unsigned long long arrPixels, int grid, int smemBytes, EmitRec * emit,
             int * emitCount, int emitCap, unsigned long long emitBase, int * err)
{
  (void)useDT; // the distance-transform median has its own pipeline (mci_dt.cu)
  if (maxNodes <= 0)
    return;
  switch (ptype)
  {
    case MCI_U8:
      launch_dil_t<uint8_t>(L, nodes, count, maxNodes, next, nextCount, nextCap, arena, arenaTop, arenaCap, in, out, dims, ball, bigHist,
                            slab, slabWords, arr, arrPixels, grid, smemBytes, emit, emitCount, emitCap, emitBase, err);
      break;
    case MCI_U16:
      launch_dil_t<uint16_t>(L, nodes, count, maxNodes, next, nextCount, nextCap, arena, arenaTop, arenaCap, in, out, dims, ball, bigHist,
                             slab, slabWords, arr, arrPixels, grid, smemBytes, emit, emitCount, emitCap, emitBase, err);
      break;
    default:
      launch_dil_t<int16_t>(L, nodes, count, maxNodes, next, nextCount, nextCap, arena, arenaTop, arenaCap, in, out, dims, ball, bigHist,
                            slab, slabWords, arr, arrPixels, grid, smemBytes, emit, emitCount, emitCap, emitBase, err);
  }
}

} // namespace mci
