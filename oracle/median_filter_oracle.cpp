// median_filter_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
//
// Restates the label smoothing step the reference's example applies to the interpolator's output
// (reference examples/mciExample.cxx:29-39):
//
//     using MedianType = itk::MedianImageFilter<ImageType, ImageType>;
//     medF->SetInput(mci->GetOutput());  medF->SetRadius(smoothingRadius);  medF->Update();
//
// itk::MedianImageFilter lives in ITK 5.4 (module ITKSmoothing), which is neither in this image nor under
// /root/reference; its published behaviour is restated here: box neighbourhood of (2r+1)^3 voxels around every
// voxel, neighbours outside the image taken from the nearest voxel inside (ZeroFluxNeumannBoundaryCondition, the
// filter's default), output = element number n/2 (0-based) of the sorted neighbourhood values
// (std::nth_element at begin + size/2).  scipy.ndimage.median_filter(size=2r+1, mode="nearest") computes the same
// function independently and pins this restatement (tests/test_median_filter.py).
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <thread>
#include <vector>

namespace
{

template <typename T>
void
median_slab(const T * in, T * out, int64_t X, int64_t Y, int64_t Z, int r, std::atomic<int64_t> & nextZ)
{
  const int        w = 2 * r + 1;
  std::vector<T>   pixels((size_t)w * w * w);
  std::vector<int64_t> xs(w), ys(w), zs(w);
  for (;;)
  {
    const int64_t z = nextZ.fetch_add(1);
    if (z >= Z)
      return;
    for (int k = 0; k < w; ++k)
      zs[k] = std::min<int64_t>(std::max<int64_t>(z + k - r, 0), Z - 1);
    for (int64_t y = 0; y < Y; ++y)
    {
      for (int k = 0; k < w; ++k)
        ys[k] = std::min<int64_t>(std::max<int64_t>(y + k - r, 0), Y - 1);
      for (int64_t x = 0; x < X; ++x)
      {
        for (int k = 0; k < w; ++k)
          xs[k] = std::min<int64_t>(std::max<int64_t>(x + k - r, 0), X - 1);
        size_t n = 0;
        for (int c = 0; c < w; ++c)
          for (int b = 0; b < w; ++b)
          {
            const T * row = in + (zs[c] * Y + ys[b]) * X;
            for (int a = 0; a < w; ++a)
              pixels[n++] = row[xs[a]];
          }
        auto medianIterator = pixels.begin() + pixels.size() / 2;
        std::nth_element(pixels.begin(), medianIterator, pixels.end());
        out[(z * Y + y) * X + x] = *medianIterator;
      }
    }
  }
}

template <typename T>
void
median_typed(const void * in, void * out, const int64_t dims[3], int r, int threads)
{
  std::atomic<int64_t>     nextZ(0);
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t)
    pool.emplace_back(median_slab<T>, (const T *)in, (T *)out, dims[0], dims[1], dims[2], r, std::ref(nextZ));
  median_slab<T>((const T *)in, (T *)out, dims[0], dims[1], dims[2], r, nextZ);
  for (std::thread & t : pool)
    t.join();
}

} // namespace

extern "C"
{
  // ptype: 0 = uint8, 1 = uint16, 2 = int16; dims = (X, Y, Z), x fastest; ITK runs this filter on all threads too
  int mci_oracle_median_image_filter(const void * in, void * out, const int64_t dims[3], int ptype, int radius, int threads)
  {
    if (!in || !out || !dims || radius < 0 || dims[0] < 1 || dims[1] < 1 || dims[2] < 1)
      return 1;
    threads = std::max(1, threads);
    switch (ptype)
    {
      case 0:
        median_typed<uint8_t>(in, out, dims, radius, threads);
        return 0;
      case 1:
        median_typed<uint16_t>(in, out, dims, radius, threads);
        return 0;
      case 2:
        median_typed<int16_t>(in, out, dims, radius, threads);
        return 0;
    }
    return 2;
  }
}
