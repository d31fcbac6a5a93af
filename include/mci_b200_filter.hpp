// mci_b200_filter.hpp -- the reference filter's public C++ interface over the C ABI of mci_b200.h.
//
// The reference class is itk::MorphologicalContourInterpolator<TImage>
// (reference include/itkMorphologicalContourInterpolator.h:64-240, "H" below; .hxx = "X").  ITK is not available in
// this build environment, so this header carries the same interface on a minimal image type instead of
// itk::Image: same method names, argument meaning, defaults, Modified()/Update() behaviour and error behaviour
// (failures surface as an exception out of Update(), as itk::ExceptionObject does, H/X: itkExceptionMacro).  With ITK
// present the same body goes into GenerateData() of the real class.
//
// Everything that touches a voxel runs on the GPU behind mci_b200.h; there is no CPU fallback: without a usable CUDA
// device Update() throws.
#ifndef MCI_B200_FILTER_HPP
#define MCI_B200_FILTER_HPP

#include <array>
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "mci_b200.h"

namespace mci_b200
{

/** What Update() throws (stands in for itk::ExceptionObject). */
class ExceptionObject : public std::runtime_error
{
public:
  ExceptionObject(int code, const std::string & what)
    : std::runtime_error(what)
    , m_Code(code)
  {}
  int
  GetCode() const
  {
    return m_Code;
  }

private:
  int m_Code;
};

namespace detail
{
inline unsigned long
NextModifiedTime()
{
  static std::atomic<unsigned long> t{ 0 };
  return ++t;
}
template <typename TPixel>
struct PixelTypeCode;
template <>
struct PixelTypeCode<unsigned char>
{
  static constexpr int value = MCI_U8;
};
template <>
struct PixelTypeCode<unsigned short>
{
  static constexpr int value = MCI_U16;
};
template <>
struct PixelTypeCode<short>
{
  static constexpr int value = MCI_I16;
};
} // namespace detail

/** Minimal stand-in for itk::Image<TPixel, 3>: x-fastest buffer, index origin 0, whole-image regions. */
template <typename TPixel, unsigned int VDimension = 3>
class Image
{
public:
  static_assert(VDimension == 3, "the B200 path covers 3-D label volumes");
  static constexpr unsigned int ImageDimension = VDimension;
  using Self = Image;
  using Pointer = std::shared_ptr<Self>;
  using ConstPointer = std::shared_ptr<const Self>;
  using PixelType = TPixel;
  using IndexValueType = std::ptrdiff_t;
  using SizeValueType = std::size_t;
  using IndexType = std::array<IndexValueType, VDimension>;
  using SizeType = std::array<SizeValueType, VDimension>;

  static Pointer
  New()
  {
    return Pointer(new Self());
  }
  void
  SetRegions(const SizeType & size)
  {
    m_Size = size;
    Modified();
  }
  const SizeType &
  GetSize() const
  {
    return m_Size;
  }
  void
  Allocate(bool initialize = false)
  {
    m_Buffer.assign(GetNumberOfPixels(), TPixel(0));
    (void)initialize; // std::vector zero-fills either way
    Modified();
  }
  SizeValueType
  GetNumberOfPixels() const
  {
    return m_Size[0] * m_Size[1] * m_Size[2];
  }
  TPixel *
  GetBufferPointer()
  {
    return m_Buffer.data();
  }
  const TPixel *
  GetBufferPointer() const
  {
    return m_Buffer.data();
  }
  SizeValueType
  ComputeOffset(const IndexType & i) const
  {
    return SizeValueType(i[0]) + m_Size[0] * (SizeValueType(i[1]) + m_Size[1] * SizeValueType(i[2]));
  }
  TPixel
  GetPixel(const IndexType & i) const
  {
    return m_Buffer[ComputeOffset(i)];
  }
  void
  SetPixel(const IndexType & i, TPixel v)
  {
    m_Buffer[ComputeOffset(i)] = v;
    Modified();
  }
  void
  FillBuffer(TPixel v)
  {
    m_Buffer.assign(GetNumberOfPixels(), v);
    Modified();
  }
  void
  Modified()
  {
    m_MTime = detail::NextModifiedTime();
  }
  unsigned long
  GetMTime() const
  {
    return m_MTime;
  }

  /** Geometry carried through unchanged (the filter works in index space, X:397): origin and the three
   *  "space directions" (direction cosines scaled by the spacing), as file formats store them. */
  using PointType = std::array<double, VDimension>;
  using SpaceDirectionsType = std::array<PointType, VDimension>;
  void
  SetOrigin(const PointType & o)
  {
    m_Origin = o;
  }
  const PointType &
  GetOrigin() const
  {
    return m_Origin;
  }
  void
  SetSpaceDirections(const SpaceDirectionsType & d)
  {
    m_SpaceDirections = d;
  }
  const SpaceDirectionsType &
  GetSpaceDirections() const
  {
    return m_SpaceDirections;
  }
  /** itk::Image::CopyInformation: geometry only, not the buffer. */
  template <typename TOther>
  void
  CopyInformation(const TOther * other)
  {
    m_Origin = other->GetOrigin();
    m_SpaceDirections = other->GetSpaceDirections();
  }

protected:
  Image() = default;

private:
  PointType           m_Origin{ { 0, 0, 0 } };
  SpaceDirectionsType m_SpaceDirections{ { { { 1, 0, 0 } }, { { 0, 1, 0 } }, { { 0, 0, 1 } } } };
  SizeType            m_Size{ { 0, 0, 0 } };
  std::vector<TPixel> m_Buffer;
  unsigned long       m_MTime{ detail::NextModifiedTime() };
};

/** Same public interface as itk::MorphologicalContourInterpolator<TImage> (H:64-224). */
template <typename TImage>
class MorphologicalContourInterpolator
{
public:
  MorphologicalContourInterpolator(const MorphologicalContourInterpolator &) = delete; // ITK_DISALLOW_COPY_AND_MOVE (H:72)
  MorphologicalContourInterpolator &
  operator=(const MorphologicalContourInterpolator &) = delete;

  using Self = MorphologicalContourInterpolator;
  using Pointer = std::shared_ptr<Self>;
  using PixelType = typename TImage::PixelType;
  using IndexValueType = typename TImage::IndexValueType;

  /** itkNewMacro(Self) (H:82) */
  static Pointer
  New()
  {
    return Pointer(new Self());
  }
  const char *
  GetNameOfClass() const
  {
    return "MorphologicalContourInterpolator";
  }
  ~MorphologicalContourInterpolator()
  {
    if (m_Ctx)
      mci_destroy(m_Ctx);
  }

#define MCI_B200_SET_GET(name, type)                                                                                           \
  void Set##name(type v)                                                                                                       \
  {                                                                                                                            \
    if (m_##name != v)                                                                                                         \
    {                                                                                                                          \
      m_##name = v;                                                                                                            \
      this->Modified();                                                                                                        \
    }                                                                                                                          \
  }                                                                                                                            \
  type Get##name() const { return m_##name; }

  /** Interpolate only this label; 0 = all labels (default) (H:85-92). */
  MCI_B200_SET_GET(Label, PixelType)
  /** Interpolate only along this axis; -1 = all axes (default) (H:94-101). */
  MCI_B200_SET_GET(Axis, int)
  /** Heuristic (default) or exhaustive alignment (H:103-113). */
  MCI_B200_SET_GET(HeuristicAlignment, bool)
  /** Distance-transform median (default, H:232) instead of repeated dilations (H:115-126). */
  MCI_B200_SET_GET(UseDistanceTransform, bool)
  /** Custom slice positions instead of auto-detection (H:128-140). */
  MCI_B200_SET_GET(UseCustomSlicePositions, bool)
  /** Ball instead of cross structuring element; also switches the 2-D components to full connectivity (H:150-165). */
  MCI_B200_SET_GET(UseBallStructuringElement, bool)
#undef MCI_B200_SET_GET
  /** Extrapolate branch extremities; default on; the reference has no getter either (H:142-147). */
  void
  SetUseExtrapolation(bool v)
  {
    if (m_UseExtrapolation != v)
    {
      m_UseExtrapolation = v;
      this->Modified();
    }
  }

  /** CUDA device the filter runs on (not in the reference). */
  void
  SetDevice(int device)
  {
    if (device != m_Device)
    {
      if (m_Ctx)
        mci_destroy(m_Ctx);
      m_Ctx = nullptr;
      m_Device = device;
      this->Modified();
    }
  }

  using SliceSetType = std::set<IndexValueType>;                      // H:175
  using LabeledSlicesType = std::unordered_map<PixelType, SliceSetType>; // H:213
  using SliceIndicesType = std::vector<LabeledSlicesType>;            // H:214

  /** Slice detection only (X:128-201): runs the detection kernels on the input, updates LabeledSliceIndices. */
  void
  DetermineSliceOrientations()
  {
    const TImage * input = this->RequireInput();
    this->RequireContext();
    int64_t dims[3];
    this->Dims(input, dims);
    this->Check(mci_upload_volume(m_Ctx, input->GetBufferPointer(), dims, detail::PixelTypeCode<PixelType>::value),
                "mci_upload_volume");
    mci_filter_options opt = this->Options();
    int64_t            n = 0;
    this->Check(mci_filter_detect(m_Ctx, &opt, &n), "mci_filter_detect");
    this->FetchLabeledSlices(n);
  }
  /** H:178-184 */
  void
  ClearLabeledSliceIndices()
  {
    m_LabeledSlices.clear();
    m_LabeledSlices.resize(TImage::ImageDimension);
    this->Modified();
  }
  /** H:188-198 */
  void
  SetLabeledSliceIndices(unsigned int axis, PixelType label, const std::vector<IndexValueType> & indices)
  {
    SliceSetType sliceSet;
    sliceSet.insert(indices.begin(), indices.end());
    m_LabeledSlices.at(axis)[label] = sliceSet;
    this->Modified();
  }
  /** H:202-206 */
  void
  SetLabeledSliceIndices(unsigned int axis, PixelType label, const SliceSetType & indices)
  {
    m_LabeledSlices.at(axis)[label] = indices;
    this->Modified();
  }
  /** H:209-212 */
  SliceSetType
  GetLabeledSliceIndices(unsigned int axis, PixelType label)
  {
    return m_LabeledSlices.at(axis)[label];
  }
  /** H:217-224 */
  SliceIndicesType
  GetLabeledSliceIndices()
  {
    return m_LabeledSlices;
  }

  // ---- ImageToImageFilter surface used by the reference's drivers (test/...Test.cxx:36-49, examples/mciExample.cxx:21-27)
  void
  SetInput(const std::shared_ptr<const TImage> & image)
  {
    if (image != m_Input)
    {
      m_Input = image;
      this->Modified();
    }
  }
  void
  SetInput(const std::shared_ptr<TImage> & image)
  {
    this->SetInput(std::shared_ptr<const TImage>(image));
  }
  const TImage *
  GetInput() const
  {
    return m_Input.get();
  }
  /** The output object; filled by Update().  Same size and type as the input. */
  std::shared_ptr<TImage>
  GetOutput()
  {
    return m_Output;
  }
  void
  Modified()
  {
    m_MTime = detail::NextModifiedTime();
  }
  unsigned long
  GetMTime() const
  {
    return m_MTime;
  }
  /** Runs GenerateData() when the filter or its input changed since the last run (ITK pipeline behaviour). */
  void
  Update()
  {
    const TImage * input = this->RequireInput();
    if (m_GeneratedAt > m_MTime && m_GeneratedAt > input->GetMTime())
      return;
    this->GenerateData();
    m_GeneratedAt = detail::NextModifiedTime();
  }
  void
  UpdateLargestPossibleRegion()
  {
    this->Update();
  }
  /** Counters and phase times of the last run (not in the reference). */
  const mci_filter_stats &
  GetLastRunStatistics() const
  {
    return m_Stats;
  }

protected:
  MorphologicalContourInterpolator()
  {
    static_assert(TImage::ImageDimension == 3, "the B200 path covers 3-D label volumes");
    m_LabeledSlices.resize(TImage::ImageDimension); // X:91-92
    m_Output = TImage::New();
  }

  /** GenerateData (X:1559-1637): zero-filled output of the input's size (AllocateOutputs, X:1541-1557), slice
   *  detection unless custom positions are set, interpolation of every (axis, label, slice pair), then the non-zero
   *  input voxels on top -- all of it inside mci_filter_run. */
  void
  GenerateData()
  {
    const TImage * input = this->RequireInput();
    this->RequireContext();
    int64_t dims[3];
    this->Dims(input, dims);
    m_Output->SetRegions(input->GetSize());
    m_Output->CopyInformation(input);
    m_Output->Allocate(true);
    mci_filter_options   opt = this->Options();
    std::vector<int64_t> custom;
    if (m_UseCustomSlicePositions)
      for (unsigned int a = 0; a < TImage::ImageDimension; ++a)
        for (const auto & kv : m_LabeledSlices[a])
          for (IndexValueType s : kv.second)
          {
            custom.push_back(int64_t(a));
            custom.push_back(int64_t(kv.first));
            custom.push_back(int64_t(s));
          }
    this->Check(mci_filter_run(m_Ctx, &opt, input->GetBufferPointer(), m_Output->GetBufferPointer(), dims,
                               detail::PixelTypeCode<PixelType>::value, custom.empty() ? nullptr : custom.data(),
                               int64_t(custom.size() / 3), &m_Stats),
                "mci_filter_run");
    if (!m_UseCustomSlicePositions)
      this->FetchLabeledSlices(m_Stats.n_annotated_slices); // DetermineSliceOrientations ran inside (X:1574-1577)
  }

private:
  const TImage *
  RequireInput() const
  {
    if (!m_Input)
      throw ExceptionObject(MCI_ERR_STATE, "MorphologicalContourInterpolator: Input is not set");
    return m_Input.get();
  }
  void
  RequireContext()
  {
    if (m_Ctx)
      return;
    int rc = mci_create(&m_Ctx, m_Device);
    if (rc != MCI_OK)
    {
      m_Ctx = nullptr;
      throw ExceptionObject(rc, "MorphologicalContourInterpolator: mci_create failed, no usable CUDA device (there is no CPU fallback)");
    }
  }
  void
  Check(int rc, const char * what)
  {
    if (rc != MCI_OK)
      throw ExceptionObject(rc, std::string("MorphologicalContourInterpolator: ") + what + ": " + mci_last_error(m_Ctx));
  }
  static void
  Dims(const TImage * image, int64_t dims[3])
  {
    for (int a = 0; a < 3; ++a)
      dims[a] = int64_t(image->GetSize()[a]);
  }
  mci_filter_options
  Options() const
  {
    mci_filter_options opt;
    mci_filter_default_options(&opt);
    opt.label = int64_t(m_Label);
    opt.axis = m_Axis;
    opt.heuristic_alignment = m_HeuristicAlignment;
    opt.use_distance_transform = m_UseDistanceTransform;
    opt.use_ball_structuring_element = m_UseBallStructuringElement;
    opt.use_custom_slice_positions = m_UseCustomSlicePositions;
    opt.use_extrapolation = m_UseExtrapolation;
    return opt;
  }
  void
  FetchLabeledSlices(int64_t n)
  {
    std::vector<int64_t> t(size_t(3 * n) + 3);
    this->Check(mci_filter_get_labeled_slices(m_Ctx, t.data()), "mci_filter_get_labeled_slices");
    m_LabeledSlices.clear();
    m_LabeledSlices.resize(TImage::ImageDimension);
    for (int64_t k = 0; k < n; ++k)
      m_LabeledSlices[size_t(t[size_t(3 * k)])][PixelType(t[size_t(3 * k + 1)])].insert(IndexValueType(t[size_t(3 * k + 2)]));
  }

  PixelType                    m_Label{ 0 };                     // X:84
  int                          m_Axis{ -1 };                     // H:230
  bool                         m_HeuristicAlignment{ true };     // H:231
  bool                         m_UseDistanceTransform{ true };   // H:232
  bool                         m_UseBallStructuringElement{ false }; // H:233
  bool                         m_UseCustomSlicePositions{ false }; // H:234
  bool                         m_UseExtrapolation{ true };       // H:235
  SliceIndicesType             m_LabeledSlices;                  // H:239
  std::shared_ptr<const TImage> m_Input;
  std::shared_ptr<TImage>      m_Output;
  unsigned long                m_MTime{ detail::NextModifiedTime() };
  unsigned long                m_GeneratedAt{ 0 };
  int                          m_Device{ 0 };
  mci_ctx *                    m_Ctx{ nullptr };
  mci_filter_stats             m_Stats{};
};

/** itk::MedianImageFilter<TInputImage, TOutputImage> as the reference's example chains it behind the interpolator
 *  (examples/mciExample.cxx:29-39): SetRadius(r) sets the same radius along every dimension; box neighbourhood,
 *  ZeroFluxNeumann boundary, element n/2 of the sorted neighbourhood.  Runs mci_median_filter_run. */
template <typename TInputImage, typename TOutputImage = TInputImage>
class MedianImageFilter
{
public:
  static_assert(std::is_same<TInputImage, TOutputImage>::value, "input and output image types must match");
  MedianImageFilter(const MedianImageFilter &) = delete;
  MedianImageFilter &
  operator=(const MedianImageFilter &) = delete;
  using Self = MedianImageFilter;
  using Pointer = std::shared_ptr<Self>;
  using PixelType = typename TInputImage::PixelType;

  static Pointer
  New()
  {
    return Pointer(new Self());
  }
  const char *
  GetNameOfClass() const
  {
    return "MedianImageFilter";
  }
  ~MedianImageFilter()
  {
    if (m_Ctx)
      mci_destroy(m_Ctx);
  }
  /** itk::BoxImageFilter::SetRadius(value); default 1 */
  void
  SetRadius(unsigned int radius)
  {
    if (radius != m_Radius)
    {
      m_Radius = radius;
      m_MTime = detail::NextModifiedTime();
    }
  }
  unsigned int
  GetRadius() const
  {
    return m_Radius;
  }
  void
  SetInput(const std::shared_ptr<const TInputImage> & image)
  {
    if (image != m_Input)
    {
      m_Input = image;
      m_MTime = detail::NextModifiedTime();
    }
  }
  void
  SetInput(const std::shared_ptr<TInputImage> & image)
  {
    this->SetInput(std::shared_ptr<const TInputImage>(image));
  }
  std::shared_ptr<TOutputImage>
  GetOutput()
  {
    return m_Output;
  }
  void
  Update()
  {
    if (!m_Input)
      throw ExceptionObject(MCI_ERR_STATE, "MedianImageFilter: Input is not set");
    if (m_GeneratedAt > m_MTime && m_GeneratedAt > m_Input->GetMTime())
      return;
    if (!m_Ctx)
    {
      int rc = mci_create(&m_Ctx, 0);
      if (rc != MCI_OK)
      {
        m_Ctx = nullptr;
        throw ExceptionObject(rc, "MedianImageFilter: mci_create failed, no usable CUDA device (there is no CPU fallback)");
      }
    }
    int64_t dims[3];
    for (int a = 0; a < 3; ++a)
      dims[a] = int64_t(m_Input->GetSize()[a]);
    m_Output->SetRegions(m_Input->GetSize());
    m_Output->CopyInformation(m_Input.get());
    m_Output->Allocate();
    int rc = mci_median_filter_run(m_Ctx, m_Input->GetBufferPointer(), m_Output->GetBufferPointer(), dims,
                                   detail::PixelTypeCode<PixelType>::value, int(m_Radius));
    if (rc != MCI_OK)
      throw ExceptionObject(rc, std::string("MedianImageFilter: mci_median_filter_run: ") + mci_last_error(m_Ctx));
    m_GeneratedAt = detail::NextModifiedTime();
  }

protected:
  MedianImageFilter() { m_Output = TOutputImage::New(); }

private:
  unsigned int                       m_Radius{ 1 };
  std::shared_ptr<const TInputImage> m_Input;
  std::shared_ptr<TOutputImage>      m_Output;
  unsigned long                      m_MTime{ detail::NextModifiedTime() };
  unsigned long                      m_GeneratedAt{ 0 };
  mci_ctx *                          m_Ctx{ nullptr };
};

} // namespace mci_b200
#endif // MCI_B200_FILTER_HPP
