// mci_kernels.cu -- kernels + launch wrappers (see mci_kernels.cuh for helpers).
#include <algorithm>
#include <cstdlib>
#include <type_traits>
#include "mci_kernels.cuh"
#include "mci_launch.h"
#include "../../include/mci_b200.h"

namespace mci
{

// ================================================================================================
// K12  slice detection + output initialisation      (DetermineSliceOrientations, X:128-201;
//      AllocateOutputs X:1541-1557 and the final overwrite X:1623-1636 folded into one pass)
// ================================================================================================
template <typename T>
__device__ __forceinline__ void
detect_voxel(const T * __restrict__ in, long long i, long long x, long long y, long long z, long long X, long long Y,
             long long Z, T val, int axisSel, int * bbox, int nl, uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2)
{
  const long long XY = X * Y;
  T nb[6];
  nb[0] = x > 0 ? in[i - 1] : T(0);
  nb[1] = x + 1 < X ? in[i + 1] : T(0);
  nb[2] = y > 0 ? in[i - X] : T(0);
  nb[3] = y + 1 < Y ? in[i + X] : T(0);
  nb[4] = z > 0 ? in[i - XY] : T(0);
  nb[5] = z + 1 < Z ? in[i + XY] : T(0);
  int cTrue = 0, cAdj = 0, axis = 0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    T p = nb[2 * a], n = nb[2 * a + 1];
    if (p == 0 && n == 0)
    {
      axis = a;
      ++cTrue;
    }
    else if (p == val && n == val)
      ++cAdj;
  }
  const unsigned li = (unsigned)(typename std::make_unsigned<T>::type)val;
  // per-label bounding box (X:148-159): test first, the table lives in L2 and rarely changes
  int c[3] = { (int)x, (int)y, (int)z };
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    int * mn = bbox + (size_t)a * nl + li;
    int * mx = bbox + (size_t)(3 + a) * nl + li;
    if (c[a] < __ldcg(mn))
      atomicMin(mn, c[a]);
    if (c[a] > __ldcg(mx))
      atomicMax(mx, c[a]);
  }
  if (cTrue == 1 && cAdj == 2 && (axisSel == -1 || axisSel == axis))
  {
    int wo = axis == 0 ? 0 : (axis == 1 ? wo1 : wo2);
    uint32_t * w = sliceBits + (size_t)li * wordsPerLabel + wo + (c[axis] >> 5);
    uint32_t bit = 1u << (c[axis] & 31);
    if (!(__ldcg(w) & bit))
      atomicOr(w, bit);
  }
}

// Non-zero chunk whose VEC voxels lie in one row and whose row length is a multiple of VEC: the y / z
// neighbours come as four aligned vector loads, the x neighbours from registers, and the table updates are
// issued once per run of equal labels instead of once per voxel.
struct ChunkSummary
{
  unsigned label; // label index + 1 of the chunk's first run, 0 = none
  int x0, x1, y, z;
  unsigned marks; // bit 0: the run marks its y slice, bit 1: its z slice
};

// per-label bounding box (X:148-159) and slice bitmap updates: test first (the tables rarely change), then atomics
__device__ __forceinline__ void
detect_update_tables(int * bbox, int nl, uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2, unsigned li, int x0, int x1, int y0,
                     int y1, int z0, int z1, unsigned markY, unsigned markZ, int ySlice, int zSlice)
{
  const int lo[3] = { x0, y0, z0 }, hi[3] = { x1, y1, z1 };
#pragma unroll
  for (int a = 0; a < 3; ++a)
  {
    int * mn = bbox + (size_t)a * nl + li;
    int * mx = bbox + (size_t)(3 + a) * nl + li;
    if (lo[a] < __ldcg(mn))
      atomicMin(mn, lo[a]);
    if (hi[a] > __ldcg(mx))
      atomicMax(mx, hi[a]);
  }
  if (markY)
  {
    uint32_t * w = sliceBits + (size_t)li * wordsPerLabel + wo1 + (ySlice >> 5);
    const uint32_t bit = 1u << (ySlice & 31);
    if (!(__ldcg(w) & bit))
      atomicOr(w, bit);
  }
  if (markZ)
  {
    uint32_t * w = sliceBits + (size_t)li * wordsPerLabel + wo2 + (zSlice >> 5);
    const uint32_t bit = 1u << (zSlice & 31);
    if (!(__ldcg(w) & bit))
      atomicOr(w, bit);
  }
}

template <typename T>
__device__ __noinline__ void
detect_chunk(const T * __restrict__ in, long long i0, uint4 q, long long X, long long Y, long long Z, int axisSel, int * bbox,
             int nl, uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2, ChunkSummary * sum)
{
  constexpr int VEC = 16 / sizeof(T);
  const long long XY = X * Y;
  long long z, y, x0;
  if (i0 + VEC <= 0xFFFFFFFFll)
  {
    // 32-bit division is several times cheaper than 64-bit and covers volumes up to 4 Gi voxels
    const unsigned iu = (unsigned)i0, xyu = (unsigned)XY, xu = (unsigned)X;
    const unsigned zu = iu / xyu, ru = iu - zu * xyu, yu = ru / xu;
    z = zu;
    y = yu;
    x0 = ru - yu * xu;
  }
  else
  {
    z = i0 / XY;
    const long long rem = i0 - z * XY;
    y = rem / X;
    x0 = rem - y * X;
  }
  T v[VEC];
  *reinterpret_cast<uint4 *>(v) = q;
  if ((X % VEC) != 0 || x0 + VEC > X)
  {
    long long x = x0, yy = y, zz = z;
    for (int e = 0; e < VEC; ++e)
    {
      if (v[e] != 0)
        detect_voxel<T>(in, i0 + e, x, yy, zz, X, Y, Z, v[e], axisSel, bbox, nl, sliceBits, wordsPerLabel, wo1, wo2);
      if (++x == X)
      {
        x = 0;
        if (++yy == Y)
        {
          yy = 0;
          ++zz;
        }
      }
    }
    return;
  }
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  const uint4 qu = y > 0 ? __ldg(reinterpret_cast<const uint4 *>(in + i0 - X)) : zero;
  const uint4 qd = y + 1 < Y ? __ldg(reinterpret_cast<const uint4 *>(in + i0 + X)) : zero;
  const uint4 qf = z > 0 ? __ldg(reinterpret_cast<const uint4 *>(in + i0 - XY)) : zero;
  const uint4 qb = z + 1 < Z ? __ldg(reinterpret_cast<const uint4 *>(in + i0 + XY)) : zero;
  const T xm = x0 > 0 ? in[i0 - 1] : T(0);
  const T xp = x0 + VEC < X ? in[i0 + VEC] : T(0);
  if (sum)
  {
    // Common case -- every labelled voxel of the chunk carries the same label: the whole 6-neighbour test
    // (X:161-197) runs as SIMD-in-register compares on the packed pixels, no per-voxel loop.
    constexpr int BITS = 8 * sizeof(T), EPW = 4 / sizeof(T);
    const unsigned c[4] = { q.x, q.y, q.z, q.w };
    const unsigned u[4] = { qu.x, qu.y, qu.z, qu.w }, d[4] = { qd.x, qd.y, qd.z, qd.w };
    const unsigned f4[4] = { qf.x, qf.y, qf.z, qf.w }, b4[4] = { qb.x, qb.y, qb.z, qb.w };
    auto eq = [](unsigned a, unsigned b) { return sizeof(T) == 2 ? __vcmpeq2(a, b) : __vcmpeq4(a, b); };
    int w0 = 0;
    while (w0 < 3 && !c[w0])
      ++w0;
    const int e0 = (__ffs(c[w0]) - 1) / BITS;
    const unsigned lab = (c[w0] >> (e0 * BITS)) & ((1u << BITS) - 1u);
    const unsigned L = sizeof(T) == 2 ? (lab | (lab << 16)) : lab * 0x01010101u;
    unsigned nz[4], my = 0, mz = 0, other = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w)
    {
      nz[w] = ~eq(c[w], 0u);
      other |= nz[w] & ~eq(c[w], L);
    }
    if (!other)
    {
      const unsigned xmW = (unsigned)(typename std::make_unsigned<T>::type)xm, xpW = (unsigned)(typename std::make_unsigned<T>::type)xp;
#pragma unroll
      for (int w = 0; w < 4; ++w)
      {
        const unsigned prevHi = w > 0 ? (c[w - 1] >> (32 - BITS)) : xmW;
        const unsigned nextLo = w < 3 ? (c[w + 1] << (32 - BITS)) : (xpW << (32 - BITS));
        const unsigned P = (c[w] << BITS) | prevHi, N = (c[w] >> BITS) | nextLo;
        const unsigned ex = eq(P, 0u) & eq(N, 0u), sx = eq(P, c[w]) & eq(N, c[w]);
        const unsigned ey = eq(u[w], 0u) & eq(d[w], 0u), sy = eq(u[w], c[w]) & eq(d[w], c[w]);
        const unsigned ez = eq(f4[w], 0u) & eq(b4[w], 0u), sz = eq(f4[w], c[w]) & eq(b4[w], c[w]);
        my |= ey & sx & sz & nz[w];
        mz |= ez & sx & sy & nz[w];
        if ((axisSel == -1 || axisSel == 0) && (ex & sy & sz & nz[w]))
        {
          unsigned m = ex & sy & sz & nz[w];
          for (int e = 0; e < EPW; ++e)
            if ((m >> (e * BITS)) & 1u)
            {
              const int cx = (int)x0 + w * EPW + e;
              uint32_t * ww = sliceBits + (size_t)lab * wordsPerLabel + (cx >> 5);
              const uint32_t bit = 1u << (cx & 31);
              if (!(__ldcg(ww) & bit))
                atomicOr(ww, bit);
            }
        }
      }
      int w1 = 3;
      while (w1 > 0 && !nz[w1])
        --w1;
      int wf = 0;
      while (wf < 3 && !nz[wf])
        ++wf;
      sum->label = lab + 1;
      sum->x0 = (int)x0 + wf * EPW + (__ffs(nz[wf]) - 1) / BITS;
      sum->x1 = (int)x0 + w1 * EPW + (31 - __clz(nz[w1])) / BITS;
      sum->y = (int)y;
      sum->z = (int)z;
      sum->marks = ((my != 0u && (axisSel == -1 || axisSel == 1)) ? 1u : 0u) | ((mz != 0u && (axisSel == -1 || axisSel == 2)) ? 2u : 0u);
      return;
    }
  }
  T up[VEC], dn[VEC], fr[VEC], bk[VEC];
  *reinterpret_cast<uint4 *>(up) = qu;
  *reinterpret_cast<uint4 *>(dn) = qd;
  *reinterpret_cast<uint4 *>(fr) = qf;
  *reinterpret_cast<uint4 *>(bk) = qb;
  int e = 0;
  while (e < VEC)
  {
    const T val = v[e];
    if (val == 0)
    {
      ++e;
      continue;
    }
    // run of equal labels [e, f)
    int f = e + 1;
    while (f < VEC && v[f] == val)
      ++f;
    unsigned mark[3] = { 0u, 0u, 0u }; // per axis: does some voxel of the run mark its slice (X:161-197)
    for (int g = e; g < f; ++g)
    {
      const T nb[6] = { g > 0 ? v[g - 1] : xm, g + 1 < VEC ? v[g + 1] : xp, up[g], dn[g], fr[g], bk[g] };
      int cTrue = 0, cAdj = 0, axis = 0;
#pragma unroll
      for (int a = 0; a < 3; ++a)
      {
        if (nb[2 * a] == 0 && nb[2 * a + 1] == 0)
        {
          axis = a;
          ++cTrue;
        }
        else if (nb[2 * a] == val && nb[2 * a + 1] == val)
          ++cAdj;
      }
      if (cTrue == 1 && cAdj == 2 && (axisSel == -1 || axisSel == axis))
      {
        if (axis == 0)
        {
          // x slices differ per voxel of the run
          const int c = (int)x0 + g;
          const unsigned li0 = (unsigned)(typename std::make_unsigned<T>::type)val;
          uint32_t * w = sliceBits + (size_t)li0 * wordsPerLabel + (c >> 5);
          const uint32_t bit = 1u << (c & 31);
          if (!(__ldcg(w) & bit))
            atomicOr(w, bit);
        }
        else
          mark[axis] = 1u;
      }
    }
    const unsigned li = (unsigned)(typename std::make_unsigned<T>::type)val;
    if (sum && !sum->label)
    {
      // first run of the chunk: handed back so that the warp can merge it with its neighbours' runs
      sum->label = li + 1;
      sum->x0 = (int)x0 + e;
      sum->x1 = (int)x0 + f - 1;
      sum->y = (int)y;
      sum->z = (int)z;
      sum->marks = mark[1] | (mark[2] << 1);
    }
    else
      detect_update_tables(bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, li, (int)x0 + e, (int)x0 + f - 1, (int)y, (int)y, (int)z,
                           (int)z, mark[1], mark[2], (int)y, (int)z);
    e = f;
  }
}

// Streaming pass: out := in with 16-byte vectors, four independent loads in flight per thread, and one
// ballot word per warp-load saying which of its 32 chunks hold labels.  Nothing else: this kernel is the
// HBM-bound part of the path (2*s*N algorithmic bytes) and must not carry the per-voxel logic's registers.
template <typename T>
__global__ void __launch_bounds__(256) k_copy_flag(const T * __restrict__ in, T * __restrict__ out, long long n,
                                                   uint32_t * __restrict__ chunkBits)
{
  constexpr int VEC = 16 / sizeof(T);
  constexpr int UNROLL = 4;
  const long long nfull = n / VEC;
  const long long stride = (long long)gridDim.x * blockDim.x; // multiple of 32: a warp-load covers one bitmap word
  const uint4 * __restrict__ in4 = reinterpret_cast<const uint4 *>(in);
  uint4 * __restrict__ out4 = reinterpret_cast<uint4 *>(out);
  const int lane = threadIdx.x & 31;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c - lane < nfull; c += UNROLL * stride)
  {
    uint4 q[UNROLL];
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
    {
      const long long ci = c + k * stride;
      q[k] = ci < nfull ? __ldcs(in4 + ci) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < UNROLL; ++k)
    {
      const long long ci = c + k * stride;
      if (ci < nfull && out)
        __stcs(out4 + ci, q[k]);
      const unsigned vote = __ballot_sync(0xFFFFFFFFu, (q[k].x | q[k].y | q[k].z | q[k].w) != 0u);
      if (lane == 0 && ci < nfull)
        chunkBits[ci >> 5] = vote;
    }
  }
}

// Per-voxel logic of DetermineSliceOrientations on the chunks flagged by k_copy_flag: one warp per bitmap word,
// one lane per 16-byte chunk, so a warp works on up to 256 consecutive voxels of a row.
template <typename T>
__global__ void __launch_bounds__(256) k_detect_marked(const T * __restrict__ in, T * __restrict__ out, long long X, long long Y,
                                                       long long Z, int axisSel, int * bbox, int nl, uint32_t * sliceBits,
                                                       int wordsPerLabel, int wo1, int wo2, const uint32_t * __restrict__ chunkBits)
{
  pdl_wait();
  constexpr int VEC = 16 / sizeof(T);
  const long long n = X * Y * Z;
  const long long nfull = n / VEC;
  const long long nWords = (nfull + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const long long gw = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
  // a warp visits words gw, gw + nw, gw + 2 nw, ...: lane i fetches the i-th of them, so the (mostly empty)
  // bitmap costs one load per 32 words instead of one dependent load per word
  for (long long wb = gw; wb < nWords; wb += 32 * nw)
  {
    const long long mine = wb + (long long)lane * nw;
    const uint32_t myBits = mine < nWords ? __ldg(chunkBits + mine) : 0u;
    unsigned nonEmpty = __ballot_sync(0xFFFFFFFFu, myBits != 0u);
    while (nonEmpty)
    {
    const int src = __ffs(nonEmpty) - 1;
    nonEmpty &= nonEmpty - 1;
    const long long w = wb + (long long)src * nw;
    const uint32_t bits = __shfl_sync(0xFFFFFFFFu, myBits, src);
    ChunkSummary sum;
    sum.label = 0;
    sum.x0 = sum.x1 = sum.y = sum.z = 0;
    sum.marks = 0;
    if ((bits >> lane) & 1u)
    {
      const long long ci = w * 32 + lane;
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(in) + ci);
      detect_chunk<T>(in, ci * VEC, q, X, Y, Z, axisSel, bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, &sum);
    }
    // lanes whose first run has the same label in the same row are merged: one table update per group
    const unsigned long long key =
      sum.label ? ((unsigned long long)sum.label | ((unsigned long long)(unsigned)sum.y << 17) | ((unsigned long long)(unsigned)sum.z << 30))
                : ~0ull; // label + 1 <= 2^16 (17 bits), y and z < 2^13
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, key);
    const int gx0 = __reduce_min_sync(grp, sum.x0), gx1 = __reduce_max_sync(grp, sum.x1);
    const unsigned gm = __reduce_or_sync(grp, sum.marks);
    if (sum.label && lane == __ffs(grp) - 1)
      detect_update_tables(bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, sum.label - 1, gx0, gx1, sum.y, sum.y, sum.z, sum.z, gm & 1u,
                           (gm >> 1) & 1u, sum.y, sum.z);
    }
  }
  // the last n % VEC voxels
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    const long long XY = X * Y;
    for (long long i = nfull * VEC; i < n; ++i)
    {
      const T val = in[i];
      if (out)
        out[i] = val;
      if (val != 0)
      {
        const long long z = i / XY, rem = i - z * XY, y = rem / X, x = rem - y * X;
        detect_voxel<T>(in, i, x, y, z, X, Y, Z, val, axisSel, bbox, nl, sliceBits, wordsPerLabel, wo1, wo2);
      }
    }
  }
}

// Compacts the per-label tables into lists the host scheduler reads back.
__global__ void
k_compact_labels(const int * __restrict__ bbox, int nl, const uint32_t * __restrict__ sliceBits, int wordsPerLabel, int wo1,
                 int wo2, int signedLabels, mci_label_info * labels, mci_labeled_slice * slices, int capSlices, int * counters)
{
  pdl_wait();
  int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= nl)
    return;
  int mnx = bbox[li], mxx = bbox[(size_t)3 * nl + li];
  if (mnx > mxx)
    return;
  long long label = signedLabels ? (long long)(short)li : (long long)li;
  int slot = atomicAdd(&counters[0], 1);
  mci_label_info L;
  L.label = label;
  for (int a = 0; a < 3; ++a)
  {
    int mn = bbox[(size_t)a * nl + li], mx = bbox[(size_t)(3 + a) * nl + li];
    L.bbox_index[a] = mn;
    L.bbox_size[a] = mx - mn + 1;
  }
  labels[slot] = L;
  const uint32_t * w = sliceBits + (size_t)li * wordsPerLabel;
  // few labels are present, so this is a handful of threads walking their bitmaps alone, and a thread's time is its
  // chain of dependent global operations: eight words are loaded before any is looked at, and the label's slots in
  // the slice list are claimed with ONE atomic (first pass: count; second pass: write)
  int total = 0;
  for (int k0 = 0; k0 < wordsPerLabel; k0 += 8)
  {
    uint32_t batch[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      batch[j] = k0 + j < wordsPerLabel ? w[k0 + j] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      total += __popc(batch[j]);
  }
  if (!total)
    return;
  int pos = atomicAdd(&counters[1], total);
  for (int k0 = 0; k0 < wordsPerLabel; k0 += 8)
  {
    uint32_t batch[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      batch[j] = k0 + j < wordsPerLabel ? w[k0 + j] : 0u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      const int k = k0 + j;
      uint32_t bits = batch[j];
      if (!bits)
        continue;
      const int axis = k >= wo2 ? 2 : (k >= wo1 ? 1 : 0);
      const int base = (k - (axis == 0 ? 0 : (axis == 1 ? wo1 : wo2))) * 32;
      while (bits)
      {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        if (pos < capSlices)
        {
          mci_labeled_slice s;
          s.axis = axis;
          s.slice = base + b;
          s.label = label;
          slices[pos] = s;
        }
        ++pos;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same streaming pass with the copy done by the TMA unit (sm_100a bulk async copies): one thread per CTA
// issues `cp.async.bulk` global -> shared (completion counted on an mbarrier) and shared -> global (bulk groups),
// COPY_STAGES tiles of 16 KB in flight per CTA, while the warps only read each tile from shared memory to build the
// "chunk holds a label" bitmap.  No voxel passes through a register on its way from `in` to `out`.
// ------------------------------------------------------------------------------------------------
#define COPY_TILE 16384 // bytes per tile: 1024 chunks of 16 bytes = 32 bitmap words
#define COPY_STAGES 4
#define COPY_THREADS 128

__device__ __forceinline__ uint32_t
smem_u32(const void * p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}

__global__ void __launch_bounds__(COPY_THREADS)
k_copy_flag_tma(const uint8_t * __restrict__ in, uint8_t * __restrict__ out, long long nTiles, uint32_t * __restrict__ chunkBits)
{
  extern __shared__ __align__(128) uint8_t tile_smem[];
  __shared__ __align__(8) unsigned long long full[COPY_STAGES];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
  {
    for (int s = 0; s < COPY_STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const long long first = blockIdx.x, step = gridDim.x;
  const long long mine = first < nTiles ? (nTiles - first + step - 1) / step : 0;
  auto issue_load = [&](long long k) {
    const int s = (int)(k % COPY_STAGES);
    const uint32_t bar = smem_u32(&full[s]), dst = smem_u32(tile_smem + (size_t)s * COPY_TILE);
    const uint8_t * src = in + (first + k * step) * (long long)COPY_TILE;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(COPY_TILE) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(COPY_TILE), "r"(bar)
                 : "memory");
  };
  if (tid == 0)
    for (long long k = 0; k < COPY_STAGES - 1 && k < mine; ++k)
      issue_load(k);
  for (long long k = 0; k < mine; ++k)
  {
    const int s = (int)(k % COPY_STAGES);
    const uint32_t parity = (uint32_t)((k / COPY_STAGES) & 1);
    // wait for the tile
    {
      const uint32_t bar = smem_u32(&full[s]);
      asm volatile("{\n"
                   ".reg .pred p;\n"
                   "WAIT_TILE:\n"
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                   "@p bra DONE_TILE;\n"
                   "bra WAIT_TILE;\n"
                   "DONE_TILE:\n"
                   "}" ::"r"(bar),
                   "r"(parity)
                   : "memory");
    }
    // bitmap: one ballot word per 32 chunks, 8 words per warp
    const uint4 * t4 = reinterpret_cast<const uint4 *>(tile_smem + (size_t)s * COPY_TILE);
    const long long tileIdx = first + k * step;
#pragma unroll
    for (int w = 0; w < (COPY_TILE / 512) / (COPY_THREADS / 32); ++w)
    {
      const int word = warp * ((COPY_TILE / 512) / (COPY_THREADS / 32)) + w;
      const uint4 q = t4[word * 32 + lane];
      const unsigned vote = __ballot_sync(0xFFFFFFFFu, (q.x | q.y | q.z | q.w) != 0u);
      if (lane == 0)
        chunkBits[tileIdx * (COPY_TILE / 512) + word] = vote;
    }
    __syncthreads(); // every warp has read the tile
    if (tid == 0)
    {
      const uint32_t src = smem_u32(tile_smem + (size_t)s * COPY_TILE);
      uint8_t * dst = out + tileIdx * (long long)COPY_TILE;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(COPY_TILE) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // the stage consumed one iteration ago is loaded next: its store must have finished READING shared memory
      if (k + COPY_STAGES - 1 < mine)
      {
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        issue_load(k + COPY_STAGES - 1);
      }
    }
  }
  if (tid == 0)
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

void
launch_detect(Launch & L, int ptype, const void * in, void * out, const long long dims[3], int axisSel, int * bbox, int nl,
              uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2, uint32_t * chunkBits)
{
  const int elt = ptype == MCI_U8 ? 1 : 2;
  const long long n = dims[0] * dims[1] * dims[2];
  const long long nchunks = n / (16 / elt);
  const long long nWords = (nchunks + 31) / 32;
  int gridA = (int)std::max<long long>(1, std::min<long long>((nchunks + 4 * 256 - 1) / (4 * 256), (long long)L.sms * 8));
  int gridB = (int)std::max<long long>(1, std::min<long long>((nWords + 7) / 8, (long long)L.sms * 8));
  L.pre("k_copy_flag");
  static const bool useTma = std::getenv("MCI_NO_TMA_COPY") == nullptr;
  long long doneVox = 0;
  // (volumes below half a gigabyte are over before the bulk pipeline has filled: C3, 134 MB, 49 us against 47 us for the
  // register path; C5, 8.6 GB, 2.79 against 3.05 ms = 0.94 against 0.86 of the measured copy bandwidth)
  if (useTma && out && n * elt >= (512ll << 20))
  {
    // full 16 KB tiles through the TMA unit; what is left (less than a tile) goes through the register path below
    const long long nTiles = (n * elt) / COPY_TILE;
    if (nTiles > 0)
    {
      static bool attr = false;
      if (!attr)
      {
        cudaFuncSetAttribute(k_copy_flag_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, COPY_STAGES * COPY_TILE);
        attr = true;
      }
      const int gridT = (int)std::min<long long>(nTiles, (long long)L.sms * 3);
      k_copy_flag_tma<<<gridT, COPY_THREADS, COPY_STAGES * COPY_TILE, L.stream>>>((const uint8_t *)in, (uint8_t *)out, nTiles, chunkBits);
      doneVox = nTiles * (COPY_TILE / elt);
    }
  }
  if (doneVox < n)
  {
    const char * inT = (const char *)in + doneVox * elt;
    char * outT = out ? (char *)out + doneVox * elt : nullptr;
    uint32_t * bitsT = chunkBits + doneVox / (16 / elt) / 32;
    const long long nT = n - doneVox;
    const long long nchT = nT / (16 / elt);
    gridA = (int)std::max<long long>(1, std::min<long long>((nchT + 4 * 256 - 1) / (4 * 256), (long long)L.sms * 8));
    switch (ptype)
    {
      case MCI_U8:
        k_copy_flag<uint8_t><<<gridA, 256, 0, L.stream>>>((const uint8_t *)inT, (uint8_t *)outT, nT, bitsT);
        break;
      case MCI_U16:
        k_copy_flag<uint16_t><<<gridA, 256, 0, L.stream>>>((const uint16_t *)inT, (uint16_t *)outT, nT, bitsT);
        break;
      default:
        k_copy_flag<int16_t><<<gridA, 256, 0, L.stream>>>((const int16_t *)inT, (int16_t *)outT, nT, bitsT);
    }
  }
  L.post("k_copy_flag");
  L.pre("k_detect_marked");
  switch (ptype)
  {
    case MCI_U8:
      launch_pdl(k_detect_marked<uint8_t>, gridB, 256, 0, L.stream, (const uint8_t *)in, (uint8_t *)out, dims[0], dims[1], dims[2], axisSel,
                                                            bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, chunkBits);
      break;
    case MCI_U16:
      launch_pdl(k_detect_marked<uint16_t>, gridB, 256, 0, L.stream, (const uint16_t *)in, (uint16_t *)out, dims[0], dims[1], dims[2],
                                                             axisSel, bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, chunkBits);
      break;
    default:
      launch_pdl(k_detect_marked<int16_t>, gridB, 256, 0, L.stream, (const int16_t *)in, (int16_t *)out, dims[0], dims[1], dims[2], axisSel,
                                                            bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, chunkBits);
  }
  L.post("k_detect_marked");
}

void
launch_compact_labels(Launch & L, const int * bbox, int nl, const uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2,
                      int signedLabels, mci_label_info * labels, mci_labeled_slice * slices, int capSlices, int * counters)
{
  L.pre("k_compact_labels");
  launch_pdl(k_compact_labels, (nl + 255) / 256, 256, 0, L.stream, bbox, nl, sliceBits, wordsPerLabel, wo1, wo2, signedLabels, labels,
                                                          slices, capSlices, counters);
  L.post("k_compact_labels");
}

// ================================================================================================
// max-combine of partial outputs (sharded runs): the write rule of X:754-763 applied voxel-wise
// ================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) k_combine_max(T * __restrict__ dst, const T * __restrict__ src, long long n)
{
  constexpr int VEC = 16 / sizeof(T);
  const long long nv = n / VEC;
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < nv; c += (long long)gridDim.x * blockDim.x)
  {
    uint4 a = reinterpret_cast<uint4 *>(dst)[c];
    const uint4 b = __ldg(reinterpret_cast<const uint4 *>(src) + c);
    T * pa = reinterpret_cast<T *>(&a);
    const T * pb = reinterpret_cast<const T *>(&b);
    bool changed = false;
#pragma unroll
    for (int e = 0; e < VEC; ++e)
      if (pb[e] > pa[e])
      {
        pa[e] = pb[e];
        changed = true;
      }
    if (changed)
      reinterpret_cast<uint4 *>(dst)[c] = a;
  }
  for (long long i = nv * VEC + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (src[i] > dst[i])
      dst[i] = src[i];
}

void
launch_combine_max(Launch & L, int ptype, void * dst, const void * src, long long n)
{
  if (n <= 0)
    return;
  const int elt = ptype == MCI_U8 ? 1 : 2;
  int grid = (int)std::max<long long>(1, std::min<long long>((n / (16 / elt) + 255) / 256, (long long)L.sms * 16));
  L.pre("k_combine_max");
  switch (ptype)
  {
    case MCI_U8:
      k_combine_max<uint8_t><<<grid, 256, 0, L.stream>>>((uint8_t *)dst, (const uint8_t *)src, n);
      break;
    case MCI_U16:
      k_combine_max<uint16_t><<<grid, 256, 0, L.stream>>>((uint16_t *)dst, (const uint16_t *)src, n);
      break;
    default:
      k_combine_max<int16_t><<<grid, 256, 0, L.stream>>>((int16_t *)dst, (const int16_t *)src, n);
  }
  L.post("k_combine_max");
}

// ================================================================================================
// K3  component matching: distinct (iId, jId) pairs of every segment  (X:1250-1272)
// ================================================================================================
__global__ void __launch_bounds__(256)
k_pairs(const SliceDev * __restrict__ slices, const SegDev * __restrict__ segs, const RowTile * __restrict__ tiles,
        const uint16_t * __restrict__ connAll, u64 * table, uint32_t tableMask, mci_pair * pairs, int capPairs, int * nPairs,
        int * err)
{
  pdl_wait();
  const RowTile tile = tiles[blockIdx.x];
  const SegDev G = segs[tile.item];
  const SliceDev A = slices[G.si], B = slices[G.sj];
  const int cw = (A.w + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int cell = warp; cell < tile.nrows * cw; cell += nwarps)
  {
    const int y = tile.row0 + cell / cw;
    const int x = (cell % cw) * 32 + lane;
    uint32_t key = 0;
    if (x < A.w)
    {
      const u64 p = (u64)y * (u64)A.w + (u64)x;
      key = ((uint32_t)connAll[A.pixOff + p] << 16) | (uint32_t)connAll[B.pixOff + p];
    }
    uint32_t prev = __shfl_up_sync(0xFFFFFFFFu, key, 1);
    if (key != 0 && (lane == 0 || key != prev))
    {
      u64 full = ((u64)(uint32_t)tile.item << 32) | key;
      uint32_t h = (uint32_t)((full * 0x9E3779B97F4A7C15ull) >> 40) & tableMask;
      for (uint32_t probe = 0; probe <= tableMask; ++probe)
      {
        u64 cur = table[h];
        if (cur == full)
          break;
        if (cur == ~0ull)
        {
          u64 old = atomicCAS(&table[h], ~0ull, full);
          if (old == ~0ull)
          {
            int pos = atomicAdd(nPairs, 1);
            if (pos < capPairs)
            {
              mci_pair q;
              q.segment = tile.item;
              q.i_id = (uint16_t)(key >> 16);
              q.j_id = (uint16_t)(key & 0xFFFFu);
              pairs[pos] = q;
            }
            else
              atomicMax(err, ERR_PAIR_TABLE);
            break;
          }
          if (old == full)
            break;
        }
        h = (h + 1) & tableMask;
        if (probe == tableMask)
          atomicMax(err, ERR_PAIR_TABLE);
      }
    }
  }
}

void
launch_pairs(Launch & L, const SliceDev * slices, const SegDev * segs, const RowTile * tiles, int ntiles, const uint16_t * conn,
             u64 * table, uint32_t tableMask, mci_pair * pairs, int capPairs, int * nPairs, int * err)
{
  if (ntiles == 0)
    return;
  L.pre("k_pairs");
  launch_pdl(k_pairs, ntiles, 256, 0, L.stream, slices, segs, tiles, conn, table, tableMask, pairs, capPairs, nPairs, err);
  L.post("k_pairs");
}

// ================================================================================================
// K6 (part)  bit-packed mask of one component over its bounding box
// ================================================================================================
__global__ void __launch_bounds__(256)
k_extract(const SliceDev * __restrict__ slices, const ExtractJob * __restrict__ jobs, const uint16_t * __restrict__ connAll,
          uint32_t * arena)
{
  pdl_wait();
  const ExtractJob J = jobs[blockIdx.x];
  const SliceDev S = slices[J.slice];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cells = J.nrows * J.m.pitch;
  for (int c = warp; c < cells; c += nwarps)
  {
    const int cell = J.row0 * J.m.pitch + c;
    const int r = cell / J.m.pitch, wi = cell % J.m.pitch;
    const int x = wi * 32 + lane;
    bool on = false;
    if (x < J.m.w)
    {
      const int u = J.m.x0 + x - S.u0, v = J.m.y0 + r - S.v0;
      on = connAll[S.pixOff + (u64)v * (u64)S.w + (u64)u] == (uint16_t)J.id;
    }
    uint32_t bits = __ballot_sync(0xFFFFFFFFu, on);
    if (lane == 0)
      arena[J.m.off + (u64)cell] = bits;
  }
}

void
launch_extract(Launch & L, const SliceDev * slices, const ExtractJob * jobs, int njobs, const uint16_t * conn, uint32_t * arena)
{
  if (njobs == 0)
    return;
  L.pre("k_extract");
  launch_pdl(k_extract, njobs, 256, 0, L.stream, slices, jobs, conn, arena);
  L.post("k_extract");
}

// ================================================================================================
// K5  Align: breadth-first translation search maximising the overlap  (X:1136-1222, Intersection X:1032-1077)
// One CTA per job.  Every translation that enters the queue is evaluated sooner or later, so all
// queued-but-unscored entries are scored in parallel (one warp each); one thread then replays the
// reference's sequential pop / compare / expand logic over those scores, which pushes the next wave.
// ================================================================================================
#define ALIGN_THREADS 1024
#define ALIGN_MASK_WORDS 12288 // shared-memory budget (words) for staging the shapes of one job
#define ALIGN_MAX_STAGED 8     // components of the other slice kept staged / addressed per job

// a shape as the scoring loop sees it: geometry + where its rows live (arena or shared memory) + row pitch there
struct AlignShape
{
  int x0, y0, w, h, words; // words = valid words per row
  int pitch;               // row stride where the data lives (odd in shared memory: conflict-free row-per-lane access)
  const uint32_t * data;
  const uint32_t * rowInfo; // staged shapes: per row first | last<<13 | single-run<<26 | nonempty<<27 (else null)
};

// overlap count |{p in A : p + t in B}| over a slice of A's rows; one lane per row, words walked with a
// running funnel shift (no per-word address arithmetic)
__device__ __forceinline__ unsigned
align_count(const AlignShape & A, const AlignShape & B, int tx, int ty, int part, int nparts, int lane)
{
  const int ylo = max(A.y0, B.y0 - ty), yhi = min(A.y0 + A.h, B.y0 + B.h - ty);
  unsigned c = 0;
  if (yhi > ylo)
  {
    const int dx0 = A.x0 + tx - B.x0; // column of B under column 0 of A's first word
    const int wlo = max(0, (-dx0) >> 5), whi = min(A.words, ((B.w - dx0 + 31) >> 5));
    if (whi > wlo)
    {
      const int rows = yhi - ylo;
      const int r0 = rows * part / nparts, r1 = rows * (part + 1) / nparts; // rows <= 16383, nparts <= 32
      const int off = dx0 >> 5, sh = dx0 & 31;
      for (int r = r0 + lane; r < r1; r += 32)
      {
        const int y = ylo + r;
        if (A.rowInfo && B.rowInfo)
        {
          // run-length shortcut: rows that are one run each overlap in an interval
          const uint32_t ia = A.rowInfo[y - A.y0], ib = B.rowInfo[y + ty - B.y0];
          if (!ia || !ib)
            continue;
          if ((ia >> 26) & (ib >> 26) & 1u)
          {
            const int a1 = A.x0 + (int)(ia & 0x1FFF), a2 = A.x0 + (int)((ia >> 13) & 0x1FFF);
            const int b1 = B.x0 + (int)(ib & 0x1FFF) - tx, b2 = B.x0 + (int)((ib >> 13) & 0x1FFF) - tx;
            c += (unsigned)max(0, min(a2, b2) - max(a1, b1) + 1);
            continue;
          }
        }
        const uint32_t * Arow = A.data + (size_t)(y - A.y0) * A.pitch;
        const uint32_t * Brow = B.data + (size_t)(y + ty - B.y0) * B.pitch;
        int bi = wlo + off;
        uint32_t lo = ((unsigned)bi < (unsigned)B.words) ? Brow[bi] : 0u;
        for (int wi = wlo; wi < whi; ++wi)
        {
          ++bi;
          const uint32_t hi = ((unsigned)bi < (unsigned)B.words) ? Brow[bi] : 0u;
          c += __popc(Arow[wi] & __funnelshift_r(lo, hi, sh));
          lo = hi;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  return c;
}

__device__ __forceinline__ AlignShape
align_shape(const MaskRef & m, const uint32_t * arena)
{
  AlignShape s;
  s.x0 = m.x0;
  s.y0 = m.y0;
  s.w = m.w;
  s.h = m.h;
  s.words = m.pitch;
  s.pitch = m.pitch;
  s.data = arena + m.off;
  s.rowInfo = nullptr;
  return s;
}

// The breadth-first search itself.  Force-inlined and instantiated twice by k_align: once with pointers
// that provably live in shared memory (LDS / STS / ATOMS), once with global scratch for huge searches.
struct AlignBfsShared
{
  int head, tail, scored, done;
};
__device__ __forceinline__ void
align_bfs(const AlignJob & J, const AlignShape & A, const AlignShape * sB, const MaskRef * __restrict__ refs,
          const uint32_t * __restrict__ arena, uint32_t * bitmap, int * qx, int * qy, unsigned * qs, int bpitch, int heuristic,
          volatile AlignBfsShared * st, unsigned * htab, int & bestx, int & besty, int & nWaves, int * err)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int qmask = J.qcap - 1;
    unsigned maxScore = 0;
    long long iter = 0;
    bestx = 0;
    besty = 0;
    if (threadIdx.x == 0)
    {
      qx[0] = 0;
      qy[0] = 0;
      qx[1] = J.indx;
      qy[1] = J.indy;
      qs[0] = qs[1] = 0u;
      bitmap[(0 - J.sy0) * bpitch + ((0 - J.sx0) >> 5)] |= 1u << ((0 - J.sx0) & 31);
      if ((unsigned)(J.indx - J.sx0) < (unsigned)J.sw && (unsigned)(J.indy - J.sy0) < (unsigned)J.sh)
        bitmap[(J.indy - J.sy0) * bpitch + ((J.indx - J.sx0) >> 5)] |= 1u << ((J.indx - J.sx0) & 31);
      st->head = 0;
      st->tail = 2;
      st->scored = 0;
      st->done = 0;
    }
    __syncthreads();
#ifdef MCI_ALIGN_PROFILE
    long long tScore = 0, tReplay = 0;
#endif
    for (;;)
    {
      const int from = st->scored, to = st->tail;
      ++nWaves;
#ifdef MCI_ALIGN_PROFILE
      const long long c0 = clock64();
#endif
      if (J.kind == JOB_EXTRAPOLATE)
      {
        // the other shape is the phantom point: score = is the point, moved back by t, inside the shape (X:239-240)
        for (int e = from + threadIdx.x; e < to; e += blockDim.x)
        {
          const int rx = J.phx - qx[e & qmask] - A.x0, ry = J.phy - qy[e & qmask] - A.y0;
          unsigned sc = 0;
          if ((unsigned)rx < (unsigned)A.w && (unsigned)ry < (unsigned)A.h)
            sc = (A.data[(size_t)ry * A.pitch + (rx >> 5)] >> (rx & 31)) & 1u;
          qs[e & qmask] = sc;
        }
      }
      else if (J.nB == 1)
      {
        // several warps share one evaluation when few translations are queued; partial counts meet in the
        // queue slot (slots are zeroed when pushed)
        const int nE = to - from, G = max(1, nwarps / max(nE, 1));
        for (int task = warp; task < nE * G; task += nwarps)
        {
          const int e = from + task / G, part = task % G;
          unsigned c = align_count(A, sB[0], qx[e & qmask], qy[e & qmask], part, G, lane);
          if (lane == 0 && c)
            atomicAdd(&qs[e & qmask], c);
        }
      }
      else
      {
        for (int e = from + warp; e < to; e += nwarps)
        {
          // iConn must intersect every listed component of jConn, else the score is 0 (X:1068-1075)
          unsigned total = 0;
          bool ok = true;
          for (int k = 0; k < J.nB && ok; ++k)
          {
            unsigned c;
            if (k < ALIGN_MAX_STAGED)
              c = align_count(A, sB[k], qx[e & qmask], qy[e & qmask], 0, 1, lane);
            else
            {
              const AlignShape Bk = align_shape(refs[J.bRef + k], arena);
              c = align_count(A, Bk, qx[e & qmask], qy[e & qmask], 0, 1, lane);
            }
            ok = c != 0;
            total += c;
          }
          if (lane == 0)
            qs[e & qmask] = ok ? total : 0u;
        }
      }
      __syncthreads();
#ifdef MCI_ALIGN_PROFILE
      const long long c1 = clock64();
      tScore += c1 - c0;
#endif
      if (warp == 0)
      {
        // The reference's sequential pop / compare / expand (X:1186-1219), replayed by one warp; all lanes carry the
        // same state.  Up to 32 queue entries are replayed at once, one per lane: the running maximum is a prefix
        // maximum, a neighbour is new when its bit is clear and no EARLIER entry of the chunk that expands is
        // adjacent to it (pushes are ordered by entry, then left / right per dimension), and the iteration count
        // before an entry is a prefix sum of the pushes.  The two `iter <=` tests of X:1198 are taken from the start
        // of the chunk; iter only grows, so they are checked against the iteration count before the chunk's last
        // entry, and the (at most two per job) chunks in which one of them flips are replayed entry by entry.
        int head = st->head, tail = to;
        while (head < to)
        {
          const int n = min(32, to - head);
          int tx = 0, ty = 0;
          unsigned sc = 0;
          if (lane < n)
          {
            const int sl = (head + lane) & qmask;
            tx = qx[sl];
            ty = qy[sl];
            sc = qs[sl];
          }
          unsigned pm = sc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
          {
            const unsigned v = __shfl_up_sync(0xFFFFFFFFu, pm, o);
            if (lane >= o)
              pm = max(pm, v);
          }
          const unsigned m = max(maxScore, pm); // maxScore once this entry has been compared (X:1191-1195)
          const bool okMin = iter <= J.minIter, okMax = iter <= J.maxIter;
          const bool ex = lane < n && (!heuristic || m == 0 || okMin || (10ull * sc > 9ull * m && okMax));
          int cx[4], cy[4];
          bool val[4];
#pragma unroll
          for (int nb = 0; nb < 4; ++nb)
          {
            cx[nb] = tx + (nb == 0 ? -1 : (nb == 1 ? 1 : 0));
            cy[nb] = ty + (nb == 2 ? -1 : (nb == 3 ? 1 : 0));
            const int rx = cx[nb] - J.sx0, ry = cy[nb] - J.sy0;
            val[nb] = ex && (unsigned)rx < (unsigned)J.sw && (unsigned)ry < (unsigned)J.sh &&
                      !((((const volatile uint32_t *)bitmap)[ry * bpitch + (rx >> 5)] >> (rx & 31)) & 1u); // volatile: the
            // bitmap may live in global scratch, where the atomicOr of an earlier chunk went to L2 past this SM's L1
          }
          // A neighbour of an expanding entry is also the neighbour of up to three other entries; it is pushed by the
          // earliest of them.  The chunk's expanding entries go into a 32 x 32 direct-mapped table keyed by the low bits
          // of their position (adjacent positions never share a slot), so each candidate looks its three other possible
          // sources up instead of comparing itself with every earlier lane.  Two entries 32 apart would share a slot:
          // then the chunk falls back to the comparison loop.
          const unsigned exMask = __ballot_sync(0xFFFFFFFFu, ex);
          int mySlot = 0;
          bool coll = false, inserted = false;
          if (ex)
          {
            const int rx = tx - J.sx0, ry = ty - J.sy0;
            mySlot = ((ry & 31) << 5) | (rx & 31);
            const unsigned key = 0x800000u | ((unsigned)lane << 18) | ((unsigned)(ry >> 5) << 9) | (unsigned)(rx >> 5);
            coll = atomicCAS(&htab[mySlot], 0u, key) != 0u;
            inserted = !coll;
          }
          __syncwarp();
          if (!__any_sync(0xFFFFFFFFu, coll))
          {
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
            {
              if (!val[nb])
                continue;
#pragma unroll
              for (int sb = 0; sb < 4; ++sb)
              {
                if (sb == (nb ^ 1))
                  continue; // that neighbour of the candidate is this lane's own entry
                const int qx_ = cx[nb] + (sb == 0 ? -1 : (sb == 1 ? 1 : 0)) - J.sx0;
                const int qy_ = cy[nb] + (sb == 2 ? -1 : (sb == 3 ? 1 : 0)) - J.sy0;
                if ((unsigned)qx_ >= (unsigned)J.sw || (unsigned)qy_ >= (unsigned)J.sh)
                  continue;
                const unsigned v = htab[((qy_ & 31) << 5) | (qx_ & 31)];
                if ((v & 0x800000u) && (v & 0x3FFFFu) == (((unsigned)(qy_ >> 5) << 9) | (unsigned)(qx_ >> 5)) &&
                    (int)((v >> 18) & 31u) < lane)
                  val[nb] = false;
              }
            }
          }
          else
          {
            for (int l = 0; l < n - 1; ++l)
            {
              const int px = __shfl_sync(0xFFFFFFFFu, tx, l), py = __shfl_sync(0xFFFFFFFFu, ty, l);
              if (((exMask >> l) & 1u) && l < lane)
              {
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
                  if (abs(cx[nb] - px) + abs(cy[nb] - py) == 1)
                    val[nb] = false;
              }
            }
          }
          __syncwarp();
          if (inserted)
            htab[mySlot] = 0u;
          const int cnt = (int)val[0] + (int)val[1] + (int)val[2] + (int)val[3];
          int incl = cnt;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1)
          {
            const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o)
              incl += v;
          }
          const int excl = incl - cnt;
          const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
          const long long iterLast = iter + __shfl_sync(0xFFFFFFFFu, excl, n - 1);
          const bool flips = heuristic && (okMin != (iterLast <= J.minIter) || okMax != (iterLast <= J.maxIter));
          if (!flips && tail + total - head <= J.qcap)
          {
            const unsigned cm = __shfl_sync(0xFFFFFFFFu, pm, 31);
            if (cm > maxScore)
            {
              const int f = __ffs(__ballot_sync(0xFFFFFFFFu, lane < n && sc == cm)) - 1; // first strictly greater
              bestx = __shfl_sync(0xFFFFFFFFu, tx, f);
              besty = __shfl_sync(0xFFFFFFFFu, ty, f);
              maxScore = cm;
            }
            int pos = tail + excl;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb)
              if (val[nb])
              {
                const int rx = cx[nb] - J.sx0, ry = cy[nb] - J.sy0;
                atomicOr(&bitmap[ry * bpitch + (rx >> 5)], 1u << (rx & 31));
                qx[pos & qmask] = cx[nb];
                qy[pos & qmask] = cy[nb];
                qs[pos & qmask] = 0u;
                ++pos;
              }
            tail += total;
            iter += total;
            head += n;
            __syncwarp();
            continue;
          }
          // entry by entry
          const int stop = head + n;
          while (head < stop)
          {
            const int tx = qx[head & qmask], ty = qy[head & qmask];
            const unsigned sc = qs[head & qmask];
            ++head;
            if (sc > maxScore)
            {
              maxScore = sc;
              bestx = tx;
              besty = ty;
            }
            // `score > maxScore * 0.9` in double (X:1198) == 10*score > 9*maxScore for integer scores below 2^32
            const bool expand = !heuristic || maxScore == 0 || iter <= J.minIter ||
                                (10ull * sc > 9ull * maxScore && iter <= J.maxIter);
            if (expand)
            {
              bool isNew = false;
              int nx = tx, ny = ty;
              if (lane < 4)
              {
                nx = tx + (lane == 0 ? -1 : (lane == 1 ? 1 : 0));
                ny = ty + (lane == 2 ? -1 : (lane == 3 ? 1 : 0));
                const int rx = nx - J.sx0, ry = ny - J.sy0;
                if ((unsigned)rx < (unsigned)J.sw && (unsigned)ry < (unsigned)J.sh)
                {
                  const uint32_t bit = 1u << (rx & 31);
                  isNew = !(atomicOr(&bitmap[ry * bpitch + (rx >> 5)], bit) & bit);
                }
              }
              const unsigned vote = __ballot_sync(0xFFFFFFFFu, isNew);
              const int nNew = __popc(vote);
              if (nNew)
              {
                if (tail + nNew - head > J.qcap)
                {
                  if (lane == 0)
                    atomicMax(err, ERR_QUEUE);
                }
                else
                {
                  if (isNew)
                  {
                    const int pos = (tail + __popc(vote & ((1u << lane) - 1u))) & qmask;
                    qx[pos] = nx;
                    qy[pos] = ny;
                    qs[pos] = 0u;
                  }
                  tail += nNew;
                }
                iter += nNew;
              }
            }
            __syncwarp();
          }
        }
        __syncwarp();
        if (lane == 0)
        {
          st->head = head;
          st->scored = to;
          st->tail = tail;
          st->done = (head == tail);
        }
      }
      __syncthreads();
#ifdef MCI_ALIGN_PROFILE
      tReplay += clock64() - c1;
#endif
      if (st->done)
        break;
    }
#ifdef MCI_ALIGN_PROFILE
    nWaves = (int)(tScore >> 4) | ((int)(tReplay >> 4) << 16); // both in units of 16 cycles
#endif
}

__global__ void __launch_bounds__(ALIGN_THREADS)
k_align(const AlignJob * __restrict__ jobs, int njobs, const MaskRef * __restrict__ refs, const uint32_t * __restrict__ arena,
        uint32_t * gscratch, int heuristic, int2 * result, int4 * dbg, int * err)
{
  pdl_wait();
  extern __shared__ uint32_t smem[];
  __shared__ AlignBfsShared s_bfs;
  __shared__ AlignShape s_A, s_B[ALIGN_MAX_STAGED];
  __shared__ unsigned s_htab[1024]; // align_bfs: positions of the expanding entries of a replay chunk
  for (int k = threadIdx.x; k < 1024; k += blockDim.x)
    s_htab[k] = 0u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int jb = blockIdx.x; jb < njobs; jb += gridDim.x)
  {
    const AlignJob J = jobs[jb];
    const long long t_begin = clock64();
    int nWaves = 0;
    const int bpitch = (J.sw + 31) >> 5;
    const int bwords = bpitch * J.sh;
    // shared memory: [staged shapes][visited bitmap][queue]; the last two move to global scratch for huge searches
    uint32_t * stage = smem;
    uint32_t * bitmap = J.useSmem ? smem + ALIGN_MASK_WORDS : gscratch + J.bitmapOff;
    uint32_t * qbuf = J.useSmem ? smem + ALIGN_MASK_WORDS + bwords : gscratch + J.queueOff;
    int * qx = (int *)qbuf;
    int * qy = qx + J.qcap;
    unsigned * qs = (unsigned *)(qy + J.qcap);
    const int qmask = J.qcap - 1;
    __syncthreads();
    // stage the shapes when they fit (rows re-pitched to an odd stride)
    if (threadIdx.x == 0)
    {
      int used = 0;
      for (int q = -1; q < J.nB && q < ALIGN_MAX_STAGED; ++q)
      {
        AlignShape s = align_shape(refs[q < 0 ? J.aRef : J.bRef + q], arena);
        const int sp = s.words | 1;
        if (used + sp * s.h + s.h <= ALIGN_MASK_WORDS)
        {
          s.pitch = -sp; // negative: "to be staged at stage + used"; fixed up below
          s.data = stage + used;
          s.rowInfo = stage + used + sp * s.h;
          used += sp * s.h + s.h;
        }
        if (q < 0)
          s_A = s;
        else
          s_B[q] = s;
      }
    }
    __syncthreads();
    for (int q = -1; q < J.nB && q < ALIGN_MAX_STAGED; ++q)
    {
      AlignShape * sp = q < 0 ? &s_A : &s_B[q];
      if (sp->pitch < 0)
      {
        const int p = -sp->pitch, wv = sp->words, total = wv * sp->h;
        const uint32_t * src = arena + refs[q < 0 ? J.aRef : J.bRef + q].off;
        uint32_t * dst = const_cast<uint32_t *>(sp->data);
        for (int k = threadIdx.x; k < total; k += blockDim.x)
        {
          const int r = k / wv, c = k - r * wv;
          dst[r * p + c] = src[k];
        }
      }
    }
    for (int k = threadIdx.x; k < bwords; k += blockDim.x)
      bitmap[k] = 0u;
    __syncthreads();
    // run summary of every staged row (one warp per row)
    for (int q = -1; q < J.nB && q < ALIGN_MAX_STAGED; ++q)
    {
      const AlignShape * sp = q < 0 ? &s_A : &s_B[q];
      if (sp->pitch >= 0)
        continue;
      const int p = -sp->pitch, wv = sp->words;
      uint32_t * info = const_cast<uint32_t *>(sp->rowInfo);
      for (int r = warp; r < sp->h; r += nwarps)
      {
        int first = 0x7FFFFFFF, last = -1, runs = 0;
        uint32_t prevTop = 0;
        for (int wb = 0; wb < wv; wb += 32)
        {
          const int wi = wb + lane;
          const uint32_t xw = wi < wv ? sp->data[r * p + wi] : 0u;
          uint32_t below = __shfl_up_sync(0xFFFFFFFFu, xw, 1);
          if (lane == 0)
            below = prevTop;
          runs += __popc(xw & ~((xw << 1) | (below >> 31)));
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
          info[r] = last >= 0 ? ((uint32_t)first | ((uint32_t)last << 13) | (runs == 1 ? (1u << 26) : 0u) | (1u << 27)) : 0u;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      if (s_A.pitch < 0)
        s_A.pitch = -s_A.pitch;
      for (int q = 0; q < J.nB && q < ALIGN_MAX_STAGED; ++q)
        if (s_B[q].pitch < 0)
          s_B[q].pitch = -s_B[q].pitch;
    }
    __syncthreads(); // pitches fixed up above
    int bestx = 0, besty = 0;
    {
      const AlignShape A = s_A;
      if (J.useSmem)
        align_bfs(J, A, s_B, refs, arena, smem + ALIGN_MASK_WORDS, (int *)(smem + ALIGN_MASK_WORDS + bwords),
                  (int *)(smem + ALIGN_MASK_WORDS + bwords) + J.qcap, (unsigned *)(smem + ALIGN_MASK_WORDS + bwords) + 2 * J.qcap,
                  bpitch, heuristic, &s_bfs, s_htab, bestx, besty, nWaves, err);
      else
        align_bfs(J, A, s_B, refs, arena, gscratch + J.bitmapOff, (int *)(gscratch + J.queueOff),
                  (int *)(gscratch + J.queueOff) + J.qcap, (unsigned *)(gscratch + J.queueOff) + 2 * J.qcap, bpitch, heuristic,
                  &s_bfs, s_htab, bestx, besty, nWaves, err);
    }
    if (threadIdx.x == 0)
    {
      result[J.out] = make_int2(bestx, besty);
      if (dbg)
        dbg[J.out] = make_int4(s_bfs.head, nWaves, (int)(clock64() - t_begin), J.kind * 1000 + J.nB);
    }
  }
}

void
launch_align(Launch & L, const AlignJob * jobs, int njobs, const MaskRef * refs, const uint32_t * arena, uint32_t * gscratch,
             int heuristic, int2 * result, int4 * dbg, int smemBytes, int * err)
{
  if (njobs == 0)
    return;
  L.ensure_max_smem(k_align, 1, L.maxSmem - 8192); // static shared memory (replay table, shape headers) takes ~5 KB
  int grid = std::min(njobs, L.sms * 8);
  // at most one job per SM: the launch lasts as long as its slowest job, and more warps shorten every scoring wave;
  // many jobs per SM: two smaller CTAs per SM keep the SM busy while one of them replays its queue -- but only if two
  // of them fit: with large search regions (visited bitmap + queue + staged shapes beyond half an SM's shared memory)
  // a second CTA would not be resident anyway, and half-width CTAs would leave half the SM idle
  const int smemPerCta = smemBytes + ALIGN_MASK_WORDS * 4 + 6 * 1024;
  const bool twoFit = 2 * smemPerCta <= L.maxSmem;
  const int threads = (njobs > L.sms && twoFit) ? ALIGN_THREADS / 2 : ALIGN_THREADS;
  L.pre("k_align");
  launch_pdl(k_align, grid, threads, smemBytes + ALIGN_MASK_WORDS * 4, L.stream, jobs, njobs, refs, arena, gscratch, heuristic,
             result, dbg, err);
  L.post("k_align");
}

// ================================================================================================
// K9  1-to-N split: synchronous multi-source geodesic growth inside the mask, ties to the
//     lowest index (Interpolate1toN, X:824-986)
// ================================================================================================
#define SPLIT_THREADS 1024
#define SPLIT_FAST_NB 4 // blobs the register-resident path holds per word

// The usual case -- mask and both plane buffers in shared memory, at most SPLIT_FAST_NB blobs, at most WPT words per
// thread: a thread keeps its words of every blob in REGISTERS for the whole growth and only the neighbour words come
// from shared memory (LDS with compile-time plane strides instead of generic loads through a pointer that may be
// global; about 40 instructions per word and step instead of about 200).  Same synchronous rule as below: every blob
// dilates from the state at the beginning of the step, a free pixel reached by several blobs goes to the lowest index.
template <int WPT, int NBMAX, bool BALL>
__device__ __forceinline__ void
split_fast(const SplitJob & J, const MaskRef & A, const int2 t, const MaskRef * __restrict__ refs, uint32_t * arena,
           uint32_t * smem, const int words, const int nB, int * s_pixels, int * err)
{
  const int pitch = A.pitch, h = A.h;
  const uint32_t * mask = smem;
  uint32_t * buf0 = smem + words;
  uint32_t * buf1 = smem + (size_t)(1 + nB) * words;
  uint32_t own[WPT][NBMAX];
  uint32_t m[WPT];
  int rr[WPT], ww[WPT];
  int c = 0;
#pragma unroll
  for (int j = 0; j < WPT; ++j)
  {
    const int idx = threadIdx.x + j * blockDim.x;
    const bool has = idx < words;
    rr[j] = has ? idx / pitch : 0;
    ww[j] = has ? idx - rr[j] * pitch : 0;
    m[j] = has ? mask[idx] : 0u;
    c += __popc(m[j]);
#pragma unroll
    for (int k = 0; k < NBMAX; ++k)
    {
      own[j][k] = 0u;
      if (k < nB && has)
      {
        // seeds: pixels of the mask whose translate lies in component k of the other slice (X:871-892)
        const MaskRef B = refs[J.bRef + k];
        own[j][k] = m[j] ? (m[j] & mask_word(arena, B, A.y0 + rr[j] + t.y, A.x0 + 32 * ww[j] + t.x)) : 0u;
        buf0[k * words + idx] = own[j][k];
      }
    }
  }
  if (c)
    atomicAdd(s_pixels, c);
  __syncthreads();
  const int maxSteps = *s_pixels + 1;
  unsigned settled = 0; // bit j: my j-th word has no free pixel left and both buffers hold its final planes
#pragma unroll
  for (int j = 0; j < WPT; ++j)
    if (threadIdx.x + j * blockDim.x >= words)
      settled |= 1u << j;
  int pending = 1;
  const uint32_t * C = buf0;
  uint32_t * N = buf1;
  for (int step = 0; step < maxSteps && pending; ++step)
  {
    int flag = 0;
#pragma unroll
    for (int j = 0; j < WPT; ++j)
    {
      if ((settled >> j) & 1u)
        continue;
      const int idx = threadIdx.x + j * blockDim.x;
      const bool up = rr[j] > 0, dn = rr[j] + 1 < h, lf = ww[j] > 0, rt = ww[j] + 1 < pitch;
      uint32_t owned = 0;
#pragma unroll
      for (int k = 0; k < NBMAX; ++k)
        owned |= own[j][k];
      const uint32_t freeBits = m[j] & ~owned;
      // neighbour words of every blob first (no branch between them: the loads of all blobs are in flight together;
      // an absent blob loads nothing, holds no pixel and therefore grows nowhere)
      uint32_t nu[NBMAX], nd[NBMAX], nl[NBMAX], nr[NBMAX];
#pragma unroll
      for (int k = 0; k < NBMAX; ++k)
      {
        const bool on = k < nB;
        const uint32_t * p = C + k * words + idx;
        nu[k] = (on && up) ? p[-pitch] : 0u;
        nd[k] = (on && dn) ? p[pitch] : 0u;
        nl[k] = (on && lf) ? p[-1] : 0u;
        nr[k] = (on && rt) ? p[1] : 0u;
        if (BALL)
        {
          nl[k] |= ((on && up && lf) ? p[-pitch - 1] : 0u) | ((on && dn && lf) ? p[pitch - 1] : 0u);
          nr[k] |= ((on && up && rt) ? p[-pitch + 1] : 0u) | ((on && dn && rt) ? p[pitch + 1] : 0u);
        }
      }
      uint32_t taken = 0;
#pragma unroll
      for (int k = 0; k < NBMAX; ++k)
      {
        const uint32_t cc = own[j][k];
        const uint32_t v = BALL ? (cc | nu[k] | nd[k]) : cc;
        uint32_t g = v | (v << 1) | (v >> 1) | (nl[k] >> 31) | (nr[k] << 31);
        if (!BALL)
          g |= nu[k] | nd[k];
        g &= freeBits & ~taken;
        own[j][k] = cc | g;
        if (k < nB)
          N[k * words + idx] = cc | g;
        taken |= g;
      }
      if (freeBits & ~taken)
        flag = 1;
      if (!freeBits)
        settled |= 1u << j;
    }
    pending = __syncthreads_or(flag);
    const uint32_t * sw = C;
    C = N;
    N = const_cast<uint32_t *>(sw);
  }
  if (pending && threadIdx.x == 0)
    atomicMax(err, ERR_SPLIT_STUCK);
#pragma unroll
  for (int j = 0; j < WPT; ++j)
  {
    const int idx = threadIdx.x + j * blockDim.x;
    if (idx < words)
    {
#pragma unroll
      for (int k = 0; k < NBMAX; ++k)
        if (k < nB)
          arena[refs[J.outRef + k].off + idx] = own[j][k];
    }
  }
}

__global__ void __launch_bounds__(SPLIT_THREADS)
k_split(const SplitJob * __restrict__ jobs, const MaskRef * __restrict__ refs, uint32_t * arena, uint32_t * gscratch,
        const int2 * __restrict__ trans, int ball, int smemWords, int fastPath, int * err)
{
  extern __shared__ uint32_t smem[];
  __shared__ int s_pixels;
  const SplitJob J = jobs[blockIdx.x];
  const MaskRef A = refs[J.aRef];
  const int2 t = trans[J.job];
  const int words = A.h * A.pitch;
  const int nB = J.nB;
  if (fastPath && nB <= SPLIT_FAST_NB && (long long)(1 + 2 * nB) * words <= smemWords && words <= 4 * (int)blockDim.x)
  {
    for (int idx = threadIdx.x; idx < words; idx += blockDim.x)
      smem[idx] = arena[A.off + idx];
    if (threadIdx.x == 0)
      s_pixels = 0;
    __syncthreads();
    const bool one = words <= (int)blockDim.x; // one word per thread (else up to four; absent words start settled)
#define SPLIT_CASE(NBMAX)                                                                        \
  if (ball)                                                                                      \
  {                                                                                              \
    if (one)                                                                                     \
      split_fast<1, NBMAX, true>(J, A, t, refs, arena, smem, words, nB, &s_pixels, err);         \
    else                                                                                         \
      split_fast<4, NBMAX, true>(J, A, t, refs, arena, smem, words, nB, &s_pixels, err);         \
  }                                                                                              \
  else                                                                                           \
  {                                                                                              \
    if (one)                                                                                     \
      split_fast<1, NBMAX, false>(J, A, t, refs, arena, smem, words, nB, &s_pixels, err);        \
    else                                                                                         \
      split_fast<4, NBMAX, false>(J, A, t, refs, arena, smem, words, nB, &s_pixels, err);        \
  }
    if (nB <= 2)
    {
      SPLIT_CASE(2)
    }
    else if (nB == 3)
    {
      SPLIT_CASE(3)
    }
    else
    {
      SPLIT_CASE(4)
    }
#undef SPLIT_CASE
    return;
  }
  // planes: mask | cur[nB] | nxt[nB]; in shared memory when they fit, else mask in the arena and the
  // ping-pong planes in the arena (final home) and global scratch
  const bool inSmem = (long long)(1 + 2 * nB) * words <= smemWords;
  const uint32_t * mask = arena + A.off;
  uint32_t * cur = inSmem ? smem + words : nullptr;
  uint32_t * nxt = inSmem ? smem + (size_t)(1 + nB) * words : nullptr;
  if (inSmem)
  {
    for (int idx = threadIdx.x; idx < words; idx += blockDim.x)
      smem[idx] = arena[A.off + idx];
    mask = smem;
  }
  auto plane = [&](bool wantCur, bool curIsFirst, int k) -> uint32_t * {
    if (inSmem)
      return (wantCur == curIsFirst ? cur : nxt) + (size_t)k * words;
    // global: "first" buffer = arena planes, "second" = scratch
    return (wantCur == curIsFirst) ? arena + refs[J.outRef + k].off : gscratch + J.scratchOff + (u64)k * words;
  };
  __syncthreads();
  // seeds: pixels of the mask whose translate lies in component k of the other slice (X:871-892)
  for (int idx = threadIdx.x; idx < words; idx += blockDim.x)
  {
    int r = idx / A.pitch, wi = idx % A.pitch;
    uint32_t m = mask[idx];
    for (int k = 0; k < nB; ++k)
    {
      const MaskRef B = refs[J.bRef + k];
      plane(true, true, k)[idx] = m ? (m & mask_word(arena, B, A.y0 + r + t.y, A.x0 + 32 * wi + t.x)) : 0u;
    }
  }
  __syncthreads();
  // growth (X:905-963): every blob dilates inside the mask at once; a free pixel reached by several blobs in
  // the same step goes to the lowest index.  A connected mask of P pixels fills in at most P steps; if free
  // pixels remain after that no seed can reach them (the reference would loop forever).
  if (threadIdx.x == 0)
    s_pixels = 0;
  __syncthreads();
  {
    int c = 0;
    for (int idx = threadIdx.x; idx < words; idx += blockDim.x)
      c += __popc(mask[idx]);
    if (c)
      atomicAdd(&s_pixels, c);
  }
  __syncthreads();
  const int maxSteps = s_pixels + 1;
  bool curIsFirst = true;
  unsigned settled = 0; // bit j: my j-th word has no free pixel left and both buffers hold its final planes
  int pending = 1;
  // the nB planes of either buffer are contiguous (stride `words`): shared memory, or the arena planes of the output
  // refs (allocated back to back) and the job's scratch planes
  uint32_t * const firstBuf = plane(true, true, 0);
  uint32_t * const secondBuf = plane(false, true, 0);
  // row / word / mask of my first word stay in registers: shapes up to blockDim.x words (the usual case, the launcher
  // sizes the CTA for it) never divide inside the loop
  const int idx0 = threadIdx.x;
  const bool has0 = idx0 < words;
  const int r0 = has0 ? idx0 / A.pitch : 0, wi0 = has0 ? idx0 - r0 * A.pitch : 0;
  const uint32_t m0 = has0 ? mask[idx0] : 0u;
  for (int step = 0; step < maxSteps && pending; ++step)
  {
    const uint32_t * C = curIsFirst ? firstBuf : secondBuf;
    uint32_t * N = curIsFirst ? secondBuf : firstBuf;
    int flag = 0;
    int j = 0;
    for (int idx = idx0; idx < words; idx += blockDim.x, ++j)
    {
      if (j < 32 && ((settled >> j) & 1u))
        continue;
      int r, wi;
      uint32_t m;
      if (j == 0)
      {
        r = r0;
        wi = wi0;
        m = m0;
      }
      else
      {
        r = idx / A.pitch;
        wi = idx - r * A.pitch;
        m = mask[idx];
      }
      uint32_t owned = 0;
      for (int k = 0; k < nB; ++k)
        owned |= C[(size_t)k * words + idx];
      const uint32_t freeBits = m & ~owned;
      uint32_t taken = 0;
      for (int k = 0; k < nB; ++k)
      {
        const uint32_t * c = C + (size_t)k * words;
        // unconditional: a step is latency bound (a few hundred of them in a row, one barrier each), so the neighbour
        // loads must not wait for the `owned` loads above
        const uint32_t g = dilate_word(c, A.h, A.pitch, r, wi, ball != 0) & freeBits & ~taken;
        N[(size_t)k * words + idx] = c[idx] | g;
        taken |= g;
      }
      if (freeBits & ~taken)
        flag = 1;
      if (!freeBits && j < 32)
        settled |= 1u << j;
    }
    pending = __syncthreads_or(flag);
    curIsFirst = !curIsFirst;
  }
  if (pending && threadIdx.x == 0)
    atomicMax(err, ERR_SPLIT_STUCK);
  // final planes -> arena
  if (inSmem || !curIsFirst)
  {
    for (int idx = threadIdx.x; idx < words; idx += blockDim.x)
      for (int k = 0; k < nB; ++k)
        arena[refs[J.outRef + k].off + idx] = plane(true, curIsFirst, k)[idx];
  }
}

void
launch_split(Launch & L, const SplitJob * jobs, int njobs, const MaskRef * refs, uint32_t * arena, uint32_t * gscratch,
             const int2 * trans, int ball, int maxWords, int * err)
{
  if (njobs == 0)
    return;
  const int smemBytes = 160 * 1024;
  L.ensure_max_smem(k_split, 2, smemBytes);
  L.pre("k_split");
  // one word per thread where the shapes allow it; fewer warps make the per-step barrier cheaper
  const int threads = std::max(128, std::min(SPLIT_THREADS, (maxWords + 31) / 32 * 32));
  static const int fastPath = getenv("MCI_SPLIT_GENERIC") ? 0 : 1; // A/B switch: the generic path for every job
  k_split<<<njobs, threads, smemBytes, L.stream>>>(jobs, refs, arena, gscratch, trans, ball, smemBytes / 4, fastPath, err);
  L.post("k_split");
}

// ================================================================================================
// root nodes of the Interpolate1to1 trees
// ================================================================================================
__global__ void
k_make_roots(const RootJob * __restrict__ roots, int nroots, MaskRef * refs, uint32_t * arena, const int2 * __restrict__ trans,
             Node * nodes)
{
  pdl_wait();
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nroots)
    return;
  const RootJob R = roots[k];
  const int2 t = trans[R.job];
  if (R.kind == JOB_EXTRAPOLATE)
  {
    // phantom point, moved against the alignment so that it lands inside the shape (X:218-250)
    MaskRef ph = refs[R.phRef];
    ph.x0 = R.phx - t.x;
    ph.y0 = R.phy - t.y;
    ph.w = ph.h = 1;
    ph.pitch = 1;
    arena[ph.off] = 1u;
    refs[R.phRef] = ph;
    Node n;
    n.A = refs[R.aRef];
    n.B = ph;
    n.tx = 0;
    n.ty = 0;
    n.i = R.posA;
    n.j = R.posB;
    n.axis = R.axis;
    n.label = R.label;
    n.slot0 = R.slot0;
    n.lo = min(R.posA, R.posB);
    nodes[R.rootOff] = n;
  }
  else if (R.kind == JOB_ONE_TO_ONE)
  {
    Node n;
    n.A = refs[R.aRef];
    n.B = refs[R.bRef];
    n.tx = t.x;
    n.ty = t.y;
    n.i = R.posA;
    n.j = R.posB;
    n.axis = R.axis;
    n.label = R.label;
    n.slot0 = R.slot0;
    n.lo = min(R.posA, R.posB);
    nodes[R.rootOff] = n;
  }
  else
  {
    for (int q = 0; q < R.nB; ++q)
    {
      Node n;
      n.A = refs[R.splitRef + q];
      n.B = refs[R.bRef + q];
      n.tx = t.x;
      n.ty = t.y;
      n.i = R.posA;
      n.j = R.posB;
      n.axis = R.axis;
      n.label = R.label;
      n.slot0 = R.slot0 < 0 ? -1 : R.slot0 + q * (abs(R.posA - R.posB) - 1);
      n.lo = min(R.posA, R.posB);
      nodes[R.rootOff + q] = n;
    }
  }
}

void
launch_make_roots(Launch & L, const RootJob * roots, int nroots, MaskRef * refs, uint32_t * arena, const int2 * trans, Node * nodes)
{
  if (nroots == 0)
    return;
  L.pre("k_make_roots");
  launch_pdl(k_make_roots, (nroots + 127) / 128, 128, 0, L.stream, roots, nroots, refs, arena, trans, nodes);
  L.post("k_make_roots");
}

} // namespace mci
