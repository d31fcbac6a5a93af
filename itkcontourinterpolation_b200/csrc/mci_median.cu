// mci_median.cu -- label smoothing after the interpolation: itk::MedianImageFilter<TImage,TImage> with
// SetRadius(r), as the reference's example applies it to the interpolator's output
// (reference examples/mciExample.cxx:29-39).
//
// Semantics (ITK MedianImageFilter): box neighbourhood (2r+1)^3, neighbours outside the image taken from the nearest
// voxel inside (ZeroFluxNeumann), output = element n/2 (0-based) of the sorted neighbourhood.
//
// One CTA per 32x8x8 output tile.  The tile and its halo are staged in shared memory once (as order-preserving 16-bit
// keys); a label volume is piecewise constant, so almost every neighbourhood is uniform and the kernel is organised
// around proving that cheaply:
//   1. tile + halo all one value (most of a sparse label volume)  -> constant store, nothing else;
//   2. otherwise per-row bit masks of value changes along x / y / z, OR-ed separably over the window, tell for every
//      output voxel whether its neighbourhood holds a change at all: if not, the output is the centre value;
//   3. the remaining voxels (those within r of a label surface) are compacted into a list and selected exactly by
//      walking the distinct neighbourhood values in ascending order (a label neighbourhood holds two or three) until
//      the cumulative count passes n/2.
// Results are staged in shared memory and stored with 16-byte vectors.  Algorithmic traffic: s*N read + s*N written.
#include <algorithm>
#include <type_traits>

#include "mci_launch.h"

namespace mci
{

namespace
{

constexpr int MED_TX = 32, MED_TY = 8, MED_TZ = 8, MED_THREADS = 256;

template <typename T>
__device__ __forceinline__ unsigned
med_key(T v)
{
  return (unsigned)v;
}
template <>
__device__ __forceinline__ unsigned
med_key<int16_t>(int16_t v)
{
  return ((unsigned)(uint16_t)v) ^ 0x8000u; // order-preserving map of signed labels onto 0..65535
}
template <typename T>
__device__ __forceinline__ T
med_unkey(unsigned k)
{
  return (T)k;
}
template <>
__device__ __forceinline__ int16_t
med_unkey<int16_t>(unsigned k)
{
  return (int16_t)(uint16_t)(k ^ 0x8000u);
}

constexpr unsigned MED_MIXED = 0xFFFFFFFFu; // tile summary: more than one value in the tile

// Streaming pre-pass: one warp per 32x8x8 tile (no halo) says whether the tile holds a single value, and which.
// A tile whose 27 neighbours (clamped at the volume border = ZeroFluxNeumann) all report the same value has a uniform
// tile + halo, so k_median3d stores the constant without reading a voxel.  Reads s*N once with 16-byte loads.
template <typename T>
__global__ void __launch_bounds__(MED_THREADS) k_median_tiles(const T * __restrict__ in, unsigned * __restrict__ summary, int X, int Y,
                                                              int Z, int tilesX, int tilesY, long long tiles)
{
  constexpr int VEC = 16 / (int)sizeof(T);   // pixels per 16-byte load
  constexpr int CPR = MED_TX / VEC;           // 16-byte chunks per tile row (4 for 16-bit pixels, 2 for 8-bit)
  const int lane = threadIdx.x & 31;
  const long long XY = (long long)X * Y;
  const bool vecOk = (X % VEC) == 0 && (reinterpret_cast<size_t>(in) & 15) == 0;
  for (long long tile = (long long)blockIdx.x * (MED_THREADS / 32) + (threadIdx.x >> 5); tile < tiles;
       tile += (long long)gridDim.x * (MED_THREADS / 32))
  {
    const int tx = (int)(tile % tilesX), ty = (int)((tile / tilesX) % tilesY), tz = (int)(tile / ((long long)tilesX * tilesY));
    const int x0 = tx * MED_TX, y0 = ty * MED_TY, z0 = tz * MED_TZ;
    const unsigned first = med_key<T>(__ldg(in + (long long)z0 * XY + (long long)y0 * X + x0));
    bool same = true;
    if (vecOk && x0 + MED_TX <= X)
    {
      // lane -> (row chunk); MED_TY * CPR chunks per z-plane
      constexpr int PER_PLANE = MED_TY * CPR;
      uint4 q[(MED_TZ * PER_PLANE + 31) / 32];
      bool have[(MED_TZ * PER_PLANE + 31) / 32];
#pragma unroll
      for (int k = 0; k < (MED_TZ * PER_PLANE + 31) / 32; ++k)
      {
        const int i = k * 32 + lane, c = i % CPR, ry = (i / CPR) % MED_TY, rz = i / PER_PLANE;
        have[k] = rz < MED_TZ && y0 + ry < Y && z0 + rz < Z;
        if (have[k])
          q[k] = __ldcs(reinterpret_cast<const uint4 *>(in + (long long)(z0 + rz) * XY + (long long)(y0 + ry) * X + x0 + c * VEC));
      }
      const T f = med_unkey<T>(first);
      unsigned pat;
      if (sizeof(T) == 2)
        pat = (unsigned)(uint16_t)f * 0x10001u;
      else
        pat = (unsigned)(uint8_t)f * 0x01010101u;
#pragma unroll
      for (int k = 0; k < (MED_TZ * PER_PLANE + 31) / 32; ++k)
        if (have[k])
          same &= ((q[k].x ^ pat) | (q[k].y ^ pat) | (q[k].z ^ pat) | (q[k].w ^ pat)) == 0u;
    }
    else
    {
      for (int i = lane; i < MED_TX * MED_TY * MED_TZ; i += 32)
      {
        const int gx = x0 + i % MED_TX, gy = y0 + (i / MED_TX) % MED_TY, gz = z0 + i / (MED_TX * MED_TY);
        if (gx < X && gy < Y && gz < Z)
          same &= med_key<T>(__ldg(in + (long long)gz * XY + (long long)gy * X + gx)) == first;
      }
    }
    same = __all_sync(0xFFFFFFFFu, same);
    if (lane == 0)
      summary[tile] = same ? first : MED_MIXED;
  }
}

constexpr int MED_DMAX = 4;  // tiles with up to this many distinct values take the bit-mask path
constexpr int MED_PADL = 8;  // a staged row holds x in [x0 - 8, x0 + 40): 48 pixels, 16-byte aligned in global memory
constexpr int MED_RW = 48;

template <typename T, int R>
__global__ void __launch_bounds__(MED_THREADS) k_median3d(const T * __restrict__ in, T * __restrict__ out, int X, int Y, int Z,
                                                          int tilesX, int tilesY, int tilesZ, const unsigned * __restrict__ summary)
{
  pdl_wait();
  constexpr int HY = MED_TY + 2 * R, HZ = MED_TZ + 2 * R;
  constexpr int NROWS = HY * HZ;            // halo rows
  constexpr int C0 = MED_PADL - R;          // column of the window's first pixel for output ox = 0
  constexpr int NWIN = (2 * R + 1) * (2 * R + 1) * (2 * R + 1);
  constexpr int KTH = NWIN / 2;             // 0-based rank of the median (std::nth_element at begin + size / 2)
  constexpr unsigned long long WMASK = (1ull << (2 * R + 1)) - 1ull;
  typedef unsigned long long u64;
  __shared__ __align__(16) uint16_t raw[NROWS * MED_RW];
  __shared__ u64 vm[(MED_DMAX - 1) * NROWS];      // vm[k][row] bit c: pixel (row, c) holds the k-th smallest value of the tile
  __shared__ u64 eX[NROWS], eY[NROWS], eZ[NROWS]; // bit c: value change between column c and its +x / +y / +z neighbour
  __shared__ u64 aX[HZ * MED_TY], aY[HZ * MED_TY], aZ[HZ * MED_TY]; // OR over the window's y range
  __shared__ u64 bX[MED_TZ * MED_TY], bM[MED_TZ * MED_TY];          // ... and over its z range
  __shared__ uint16_t res[MED_TZ * MED_TY * MED_TX];
  __shared__ uint16_t list[MED_TZ * MED_TY * MED_TX];
  __shared__ int listCount;
  __shared__ int sLo, sHi, sNext[MED_DMAX];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = blockIdx.x, ty = blockIdx.y, tz = blockIdx.z; // grid = tiles along x, y, z
  const long long tile = ((long long)tz * tilesY + ty) * tilesX + tx;
  const int x0 = tx * MED_TX, y0 = ty * MED_TY, z0 = tz * MED_TZ;
  const long long XY = (long long)X * Y;
  if (tid == 0)
  {
    listCount = 0;
    sLo = 0x10000;
    sHi = -1;
    for (int k = 0; k < MED_DMAX; ++k)
      sNext[k] = 0x10000;
  }
  const int ox = lane, orow0 = warp; // output voxel (ox, row) with row = oy + MED_TY * oz, rows orow0, orow0 + 8, ...
  constexpr int VEC = 16 / (int)sizeof(T);
  const bool vecOk = (X % VEC) == 0 && x0 + MED_TX <= X && (reinterpret_cast<size_t>(out) & 15) == 0;

  // stores one value to the whole tile
  auto storeConstant = [&](unsigned key) {
    const T v = med_unkey<T>(key);
    if (vecOk)
    {
      T tmp[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        tmp[k] = v;
      constexpr int VPR = MED_TX / VEC;
      for (int i = tid; i < MED_TY * MED_TZ * VPR; i += MED_THREADS)
      {
        const int c = i % VPR, row = i / VPR, gy = y0 + row % MED_TY, gz = z0 + row / MED_TY;
        if (gy < Y && gz < Z)
          __stcs(reinterpret_cast<uint4 *>(out + (long long)gz * XY + (long long)gy * X + x0 + c * VEC), *reinterpret_cast<const uint4 *>(tmp));
      }
    }
    else
      for (int row = orow0; row < MED_TY * MED_TZ; row += MED_THREADS / 32)
      {
        const int gy = y0 + (row % MED_TY), gz = z0 + (row / MED_TY), gx = x0 + ox;
        if (gx < X && gy < Y && gz < Z)
          out[(long long)gz * XY + (long long)gy * X + gx] = v;
      }
  };

  // ---- 0. the 27 tile summaries around this tile (halo <= 3 < tile extent): all one value -> constant store
  {
    const unsigned mine = summary[tile];
    bool same = mine != MED_MIXED;
    if (tid < 27)
    {
      const int nx = min(max(tx + tid % 3 - 1, 0), tilesX - 1), ny = min(max(ty + (tid / 3) % 3 - 1, 0), tilesY - 1),
                nz = min(max(tz + tid / 9 - 1, 0), tilesZ - 1);
      same = summary[((long long)nz * tilesY + ny) * tilesX + nx] == mine;
    }
    if (__syncthreads_and(same) && mine != MED_MIXED)
    {
      storeConstant(mine);
      return;
    }
  }

  // ---- 1. stage 48-pixel rows of tile + halo as 16-bit keys (clamped coordinates = ZeroFluxNeumann)
  if (sizeof(T) == 2 && (X % 8) == 0 && (reinterpret_cast<size_t>(in) & 15) == 0)
  {
    // 16-bit pixels, rows 16-byte aligned: six 16-byte loads per row; a chunk beyond the volume's x range repeats
    // the border pixel
    constexpr int CPR = MED_RW / 8;
    for (int i = tid; i < NROWS * CPR; i += MED_THREADS)
    {
      const int c = i % CPR, row = i / CPR, hy = row % HY, hz = row / HY;
      const int gy = min(max(y0 + hy - R, 0), Y - 1), gz = min(max(z0 + hz - R, 0), Z - 1);
      const int xs = x0 - MED_PADL + 8 * c;
      const T * rowp = in + (long long)gz * XY + (long long)gy * X;
      uint4 q = __ldg(reinterpret_cast<const uint4 *>(rowp + min(max(xs, 0), X - 8)));
      if (xs < 0)
      {
        const unsigned e = (q.x & 0xFFFFu) * 0x10001u;
        q = make_uint4(e, e, e, e);
      }
      else if (xs >= X)
      {
        const unsigned e = (q.w >> 16) * 0x10001u;
        q = make_uint4(e, e, e, e);
      }
      if (std::is_same<T, int16_t>::value)
      {
        q.x ^= 0x80008000u;
        q.y ^= 0x80008000u;
        q.z ^= 0x80008000u;
        q.w ^= 0x80008000u;
      }
      reinterpret_cast<uint4 *>(raw)[i] = q;
    }
  }
  else
  {
    for (int i = tid; i < NROWS * MED_RW; i += MED_THREADS)
    {
      const int c = i % MED_RW, row = i / MED_RW, hy = row % HY, hz = row / HY;
      const int gx = min(max(x0 - MED_PADL + c, 0), X - 1), gy = min(max(y0 + hy - R, 0), Y - 1), gz = min(max(z0 + hz - R, 0), Z - 1);
      raw[i] = (uint16_t)med_key<T>(__ldg(in + (long long)gz * XY + (long long)gy * X + gx));
    }
  }
  __syncthreads();

  // ---- 2. the distinct values of the staged block in ascending order, as long as there are at most MED_DMAX
  unsigned vals[MED_DMAX], vHi;
  int nvals;
  {
    constexpr int NW = NROWS * MED_RW / 2; // 32-bit words, two pixels each
    const unsigned * rw = reinterpret_cast<const unsigned *>(raw);
    int lo = 0x10000, hi = -1;
    for (int i = tid; i < NW; i += MED_THREADS)
    {
      const unsigned w = rw[i];
      const int a = (int)(w & 0xFFFFu), b = (int)(w >> 16);
      lo = min(lo, min(a, b));
      hi = max(hi, max(a, b));
    }
    lo = __reduce_min_sync(0xFFFFFFFFu, lo);
    hi = __reduce_max_sync(0xFFFFFFFFu, hi);
    if (lane == 0)
    {
      atomicMin(&sLo, lo);
      atomicMax(&sHi, hi);
    }
    __syncthreads();
    lo = sLo;
    hi = sHi;
    if (lo == hi)
    {
      storeConstant((unsigned)lo); // the block inside the 27 tiles is uniform after all
      return;
    }
    vals[0] = (unsigned)lo;
    nvals = 1;
    vHi = (unsigned)hi;
    // values strictly between the last one found and the maximum, smallest first; round r fills vals[r]
    bool done = false;
#pragma unroll
    for (int r = 1; r < MED_DMAX; ++r)
      if (!done)
      {
        const int last = (int)vals[r - 1];
        int nxt = 0x10000;
        for (int i = tid; i < NW; i += MED_THREADS)
        {
          const unsigned w = rw[i];
          const int a = (int)(w & 0xFFFFu), b = (int)(w >> 16);
          if (a > last && a < hi)
            nxt = min(nxt, a);
          if (b > last && b < hi)
            nxt = min(nxt, b);
        }
        nxt = __reduce_min_sync(0xFFFFFFFFu, nxt);
        if (lane == 0 && nxt < 0x10000)
          atomicMin(&sNext[r], nxt);
        __syncthreads();
        nxt = sNext[r];
        if (nxt == 0x10000)
          done = true;
        else if (r == MED_DMAX - 1)
        {
          nvals = MED_DMAX + 1; // a fifth value: generic path
          done = true;
        }
        else
        {
          vals[r] = (unsigned)nxt;
          nvals = r + 1;
        }
      }
    if (nvals <= MED_DMAX - 1)
    {
#pragma unroll
      for (int k = 1; k < MED_DMAX; ++k)
        if (k == nvals)
          vals[k] = vHi;
      ++nvals;
    }
  }

  if (nvals <= MED_DMAX)
  {
    // ---- 3a. value masks by warp ballot (one warp per row; the largest value needs no mask) ...
    for (int row = warp; row < NROWS; row += MED_THREADS / 32)
    {
      const unsigned v0 = raw[row * MED_RW + lane];
      const unsigned v1 = lane < MED_RW - 32 ? raw[row * MED_RW + 32 + lane] : 0x10000u;
#pragma unroll
      for (int k = 0; k < MED_DMAX - 1; ++k)
        if (k < nvals - 1)
        {
          const unsigned b0 = __ballot_sync(0xFFFFFFFFu, v0 == vals[k]), b1 = __ballot_sync(0xFFFFFFFFu, v1 == vals[k]);
          if (lane == 0)
            vm[k * NROWS + row] = (u64)b0 | ((u64)b1 << 32);
        }
    }
    __syncthreads();
    // ... and the value-change masks from them: pixels c and c' differ iff some value mask differs between them
    for (int row = tid; row < NROWS; row += MED_THREADS)
    {
      const int hy = row % HY, hz = row / HY;
      u64 mx = 0ull, my = 0ull, mz = 0ull;
#pragma unroll
      for (int k = 0; k < MED_DMAX - 1; ++k)
        if (k < nvals - 1)
        {
          const u64 m = vm[k * NROWS + row];
          mx |= m ^ (m >> 1);
          if (hy + 1 < HY)
            my |= m ^ vm[k * NROWS + row + 1];
          if (hz + 1 < HZ)
            mz |= m ^ vm[k * NROWS + row + HY];
        }
      eX[row] = mx; // bit 47 compares with a pixel that is not staged; windows never reach it
      eY[row] = my;
      eZ[row] = mz;
    }
  }
  else
  {
    // ---- 3b. many distinct values (not a label volume): value-change masks pixel by pixel
    for (int row = tid; row < NROWS; row += MED_THREADS)
    {
      const int hy = row % HY, hz = row / HY;
      const uint16_t * p = raw + row * MED_RW;
      const uint16_t * py = hy + 1 < HY ? p + MED_RW : p;
      const uint16_t * pz = hz + 1 < HZ ? p + HY * MED_RW : p;
      u64 mx = 0ull, my = 0ull, mz = 0ull;
      unsigned v = p[0];
#pragma unroll 4
      for (int c = 0; c < MED_RW; ++c)
      {
        const unsigned nx = c + 1 < MED_RW ? p[c + 1] : v;
        mx |= (u64)(v != nx) << c;
        my |= (u64)(v != py[c]) << c;
        mz |= (u64)(v != pz[c]) << c;
        v = nx;
      }
      eX[row] = mx;
      eY[row] = my;
      eZ[row] = mz;
    }
  }
  __syncthreads();
  // OR over the y range of the window: x- and z-changes over [oy, oy + 2R], y-changes over [oy, oy + 2R - 1]
  for (int i = tid; i < HZ * MED_TY; i += MED_THREADS)
  {
    const int oy = i % MED_TY, hz = i / MED_TY;
    u64 sx = 0ull, sy = 0ull, sz = 0ull;
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d)
    {
      sx |= eX[hz * HY + oy + d];
      sz |= eZ[hz * HY + oy + d];
      if (d < 2 * R)
        sy |= eY[hz * HY + oy + d];
    }
    aX[i] = sx;
    aY[i] = sy;
    aZ[i] = sz;
  }
  __syncthreads();
  // ... and over its z range: x- and y-changes over [oz, oz + 2R], z-changes over [oz, oz + 2R - 1]
  for (int i = tid; i < MED_TZ * MED_TY; i += MED_THREADS)
  {
    const int oy = i % MED_TY, oz = i / MED_TY;
    u64 sx = 0ull, sm = 0ull;
#pragma unroll
    for (int d = 0; d <= 2 * R; ++d)
    {
      sx |= aX[(oz + d) * MED_TY + oy];
      sm |= aY[(oz + d) * MED_TY + oy];
      if (d < 2 * R)
        sm |= aZ[(oz + d) * MED_TY + oy];
    }
    bX[i] = sx;
    bM[i] = sm;
  }
  __syncthreads();

  // ---- 4. uniform neighbourhoods take the centre value; the others go on the list
  for (int row = orow0; row < MED_TY * MED_TZ; row += MED_THREADS / 32)
  {
    const int oy = row % MED_TY, oz = row / MED_TY;
    // x-changes between c and c + 1 for c in [ox + C0, ox + C0 + 2R - 1]; y/z-changes at c in [ox + C0, ox + C0 + 2R]
    const u64 hit = ((bX[row] >> (ox + C0)) & (WMASK >> 1)) | ((bM[row] >> (ox + C0)) & WMASK);
    const int o = row * MED_TX + ox;
    if (hit == 0ull)
      res[o] = raw[((oz + R) * HY + oy + R) * MED_RW + ox + MED_PADL];
    else if (x0 + ox < X && y0 + oy < Y && z0 + oz < Z)
      list[atomicAdd(&listCount, 1)] = (uint16_t)o;
  }
  __syncthreads();

  // ---- 5. exact selection for the listed voxels
  const int nList = listCount;
  if (nvals <= MED_DMAX)
  {
    // counts of the tile's values inside the window by popcount of the value masks, smallest value first, until the
    // cumulative count passes n/2
    for (int e = tid; e < nList; e += MED_THREADS)
    {
      const int o = list[e];
      const int vx = o % MED_TX, vy = (o / MED_TX) % MED_TY, vz = o / (MED_TX * MED_TY);
      const int sh = vx + C0;
      unsigned answer = vHi;
      int total = 0;
#pragma unroll
      for (int k = 0; k < MED_DMAX - 1; ++k)
        if (k < nvals - 1)
        {
          const u64 * m = vm + k * NROWS + vz * HY + vy;
          int c = 0;
#pragma unroll
          for (int dz = 0; dz <= 2 * R; ++dz)
#pragma unroll
            for (int dy = 0; dy <= 2 * R; ++dy)
              c += __popc((unsigned)((m[dz * HY + dy] >> sh) & WMASK));
          total += c;
          if (total > KTH)
          {
            answer = vals[k];
            break;
          }
        }
      res[o] = (uint16_t)answer;
    }
  }
  else
  {
    // distinct values of the neighbourhood in ascending order until the cumulative count passes n/2
    for (int e = tid; e < nList; e += MED_THREADS)
    {
      const int o = list[e];
      const int vx = o % MED_TX, vy = (o / MED_TX) % MED_TY, vz = o / (MED_TX * MED_TY);
      const uint16_t * base = raw + (vz * HY + vy) * MED_RW + vx + C0;
      int cur = -1, total = 0, m;
      for (;;)
      {
        m = 0x10000;
        int c = 0;
        for (int dz = 0; dz <= 2 * R; ++dz)
          for (int dy = 0; dy <= 2 * R; ++dy)
          {
            const uint16_t * p = base + (dz * HY + dy) * MED_RW;
#pragma unroll
            for (int dx = 0; dx <= 2 * R; ++dx)
            {
              const int v = p[dx];
              if (v > cur)
              {
                c = v < m ? 1 : c + (v == m);
                m = min(m, v);
              }
            }
          }
        total += c;
        if (total > KTH)
          break;
        cur = m;
      }
      res[o] = (uint16_t)m;
    }
  }
  __syncthreads();

  // ---- 6. store the tile: 16-byte vectors when rows are aligned and the tile is inside the volume
  if (vecOk)
  {
    constexpr int VPR = MED_TX / VEC; // vectors per row
    for (int i = tid; i < MED_TY * MED_TZ * VPR; i += MED_THREADS)
    {
      const int v = i % VPR, row = i / VPR, gy = y0 + row % MED_TY, gz = z0 + row / MED_TY;
      if (gy >= Y || gz >= Z)
        continue;
      T tmp[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k)
        tmp[k] = med_unkey<T>(res[row * MED_TX + v * VEC + k]);
      *reinterpret_cast<uint4 *>(out + (long long)gz * XY + (long long)gy * X + x0 + v * VEC) = *reinterpret_cast<const uint4 *>(tmp);
    }
  }
  else
  {
    for (int row = orow0; row < MED_TY * MED_TZ; row += MED_THREADS / 32)
    {
      const int gy = y0 + row % MED_TY, gz = z0 + row / MED_TY, gx = x0 + ox;
      if (gx < X && gy < Y && gz < Z)
        out[(long long)gz * XY + (long long)gy * X + gx] = med_unkey<T>(res[row * MED_TX + ox]);
    }
  }
}

template <typename T>
void
launch_median_typed(Launch & L, const T * in, T * out, const long long dims[3], int radius, unsigned * summary)
{
  const int X = (int)dims[0], Y = (int)dims[1], Z = (int)dims[2];
  const int tilesX = (X + MED_TX - 1) / MED_TX, tilesY = (Y + MED_TY - 1) / MED_TY, tilesZ = (Z + MED_TZ - 1) / MED_TZ;
  const long long tiles = (long long)tilesX * tilesY * tilesZ;
  const int gridA = (int)std::max<long long>(1, std::min<long long>((tiles + 7) / 8, (long long)L.sms * 16));
  L.pre("k_median_tiles");
  k_median_tiles<T><<<gridA, MED_THREADS, 0, L.stream>>>(in, summary, X, Y, Z, tilesX, tilesY, tiles);
  L.post("k_median_tiles");
  const dim3 grid3((unsigned)tilesX, (unsigned)tilesY, (unsigned)tilesZ); // each volume dimension is at most 8192: <= 1024 tiles
  L.pre("k_median3d");
  switch (radius)
  {
    case 1:
      launch_pdl(k_median3d<T, 1>, grid3, MED_THREADS, 0, L.stream, in, out, X, Y, Z, tilesX, tilesY, tilesZ, summary);
      break;
    case 2:
      launch_pdl(k_median3d<T, 2>, grid3, MED_THREADS, 0, L.stream, in, out, X, Y, Z, tilesX, tilesY, tilesZ, summary);
      break;
    default:
      launch_pdl(k_median3d<T, 3>, grid3, MED_THREADS, 0, L.stream, in, out, X, Y, Z, tilesX, tilesY, tilesZ, summary);
  }
  L.post("k_median3d");
}

} // namespace

long long
median_tile_count(const long long dims[3])
{
  return ((dims[0] + MED_TX - 1) / MED_TX) * ((dims[1] + MED_TY - 1) / MED_TY) * ((dims[2] + MED_TZ - 1) / MED_TZ);
}

// radius in [1, 3]; in != out; summary: median_tile_count(dims) words of scratch
void
launch_median3d(Launch & L, int ptype, const void * in, void * out, const long long dims[3], int radius, unsigned * summary)
{
  switch (ptype)
  {
    case MCI_U8:
      launch_median_typed<uint8_t>(L, (const uint8_t *)in, (uint8_t *)out, dims, radius, summary);
      break;
    case MCI_U16:
      launch_median_typed<uint16_t>(L, (const uint16_t *)in, (uint16_t *)out, dims, radius, summary);
      break;
    default:
      launch_median_typed<int16_t>(L, (const int16_t *)in, (int16_t *)out, dims, radius, summary);
  }
}

} // namespace mci
