// mci_kernels.cuh -- hand-written sm_100a kernels of the slice-pair interpolation path.
// Every kernel is batched over all slices / segments / jobs / nodes in flight.
// Reference citations "X:n" are lines of include/itkMorphologicalContourInterpolator.hxx.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mci_types.h"

namespace mci
{

#define MCI_NOLABEL 0xFFFFFFFFu

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------

// 32 bits of row y of mask m starting at absolute slice column X (zero outside the mask).
__device__ __forceinline__ uint32_t
mask_word(const uint32_t * __restrict__ arena, const MaskRef & m, int y, int X)
{
  int ry = y - m.y0;
  if ((unsigned)ry >= (unsigned)m.h)
    return 0u;
  int rx = X - m.x0;
  int wi = rx >> 5; // floor
  int sh = rx & 31;
  const uint32_t * row = arena + m.off + (u64)ry * (u64)m.pitch;
  uint32_t lo = ((unsigned)wi < (unsigned)m.pitch) ? row[wi] : 0u;
  uint32_t hi = ((unsigned)(wi + 1) < (unsigned)m.pitch) ? row[wi + 1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

// same, with `data` pointing at the first word of the mask (global or shared memory)
__device__ __forceinline__ uint32_t
mask_word_p(const uint32_t * data, const MaskRef & m, int y, int X)
{
  int ry = y - m.y0;
  if ((unsigned)ry >= (unsigned)m.h)
    return 0u;
  int rx = X - m.x0;
  int wi = rx >> 5;
  int sh = rx & 31;
  const uint32_t * row = data + (size_t)ry * m.pitch;
  uint32_t lo = ((unsigned)wi < (unsigned)m.pitch) ? row[wi] : 0u;
  uint32_t hi = ((unsigned)(wi + 1) < (unsigned)m.pitch) ? row[wi + 1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

__device__ __forceinline__ bool
mask_bit(const uint32_t * __restrict__ arena, const MaskRef & m, int x, int y)
{
  int rx = x - m.x0, ry = y - m.y0;
  if ((unsigned)rx >= (unsigned)m.w || (unsigned)ry >= (unsigned)m.h)
    return false;
  return (arena[m.off + (u64)ry * (u64)m.pitch + (rx >> 5)] >> (rx & 31)) & 1u;
}

// radius-1 dilation of the word at (row r, word wi) of a bit image of h rows x pitch words.
// cross: centre + 4 face neighbours; ball: full 3x3 box (ITK radius-1 ball).  Outside = background.
__device__ __forceinline__ uint32_t
hdil(const uint32_t * row, int pitch, int wi)
{
  uint32_t c = row[wi];
  uint32_t l = wi > 0 ? row[wi - 1] : 0u;
  uint32_t r = wi + 1 < pitch ? row[wi + 1] : 0u;
  return c | (c << 1) | (l >> 31) | (c >> 1) | (r << 31);
}
__device__ __forceinline__ uint32_t
dilate_word(const uint32_t * img, int h, int pitch, int r, int wi, bool ball)
{
  const uint32_t * row = img + (size_t)r * pitch;
  uint32_t v = hdil(row, pitch, wi);
  if (ball)
  {
    if (r > 0)
      v |= hdil(row - pitch, pitch, wi);
    if (r + 1 < h)
      v |= hdil(row + pitch, pitch, wi);
  }
  else
  {
    if (r > 0)
      v |= row[wi - pitch];
    if (r + 1 < h)
      v |= row[wi + pitch];
  }
  return v;
}

// nearest set bit of a bit row to column x, as squared distance, searching only while it can beat lim
__device__ __forceinline__ int
row_nearest_sq(const uint32_t * row, int pitch, int x, int lim)
{
  const int w0 = x >> 5, b = x & 31;
  int best = lim;
  {
    uint32_t m = row[w0] & (0xFFFFFFFFu >> (31 - b));
    int wl = w0;
    for (;;)
    {
      if (m)
      {
        int dx = x - (wl * 32 + 31 - __clz(m));
        best = min(best, dx * dx);
        break;
      }
      if (--wl < 0)
        break;
      int dmin = x - (wl * 32 + 31);
      if (dmin * dmin >= best)
        break;
      m = row[wl];
    }
  }
  {
    uint32_t m = row[w0] & (0xFFFFFFFEu << b);
    int wr = w0;
    for (;;)
    {
      if (m)
      {
        int dx = (wr * 32 + __ffs(m) - 1) - x;
        best = min(best, dx * dx);
        break;
      }
      if (++wr >= pitch)
        break;
      int dmin = wr * 32 - x;
      if (dmin * dmin >= best)
        break;
      m = row[wr];
    }
  }
  return best;
}

template <typename T>
struct PixTraits;
template <>
struct PixTraits<uint8_t>
{
  static const int kMaxComp = 255;
  static const int kBins = 256;
};
template <>
struct PixTraits<uint16_t>
{
  static const int kMaxComp = 65535;
  static const int kBins = 65536;
};
template <>
struct PixTraits<int16_t>
{
  static const int kMaxComp = 32767;
  static const int kBins = 32768;
};

// out = max(out, label) on a voxel (write rule of X:754-763), safe under concurrent writers.
__device__ __forceinline__ void
atomic_max_pixel(uint16_t * p, uint16_t v)
{
  unsigned short old = *p;
  while (old < v)
  {
    unsigned short prev = atomicCAS(reinterpret_cast<unsigned short *>(p), old, v);
    if (prev == old)
      break;
    old = prev;
  }
}
__device__ __forceinline__ void
atomic_max_pixel(int16_t * p, int16_t v)
{
  unsigned short * q = reinterpret_cast<unsigned short *>(p);
  unsigned short old = *q;
  while ((short)old < v)
  {
    unsigned short prev = atomicCAS(q, old, (unsigned short)v);
    if (prev == old)
      break;
    old = prev;
  }
}
__device__ __forceinline__ void
atomic_max_pixel(uint8_t * p, uint8_t v)
{
  size_t a = reinterpret_cast<size_t>(p);
  unsigned * w = reinterpret_cast<unsigned *>(a & ~size_t(3));
  unsigned sh = unsigned(a & 3) * 8;
  unsigned old = *w;
  while (((old >> sh) & 0xFFu) < v)
  {
    unsigned nv = (old & ~(0xFFu << sh)) | (unsigned(v) << sh);
    unsigned prev = atomicCAS(w, old, nv);
    if (prev == old)
      break;
    old = prev;
  }
}

} // namespace mci
