// Test driver for the C++ filter class of include/mci_b200_filter.hpp, shaped like the reference's own driver
// (reference test/itkMorphologicalContourInterpolationTest.cxx:25-140): read an image, set the algorithm / axis /
// label on the filter, write the filter's output.  Images are raw x-fastest voxel files here (ITK's readers are not
// available): the pytest side converts fixtures and compares the written output with the oracle.
//
//   mci_filter_test inputImage.nrrd outputImage.nrrd [algorithm B|C|T] [axis] [label]
//                                the reference driver's own command line on NRRD files: pixel type chosen as it
//                                chooses it (3-D short / ushort -> Image<short,3>), output written with compression;
//                                the written file is byte-compatible with ITK's writer (include/mci_b200_io.hpp)
//   mci_filter_test inputRaw outputRaw X Y Z uchar|ushort|short [algorithm B|C|T|D] [axis] [label]
//   mci_filter_test --example inputRaw outputRaw X Y Z uchar|ushort|short [smoothingRadius=2]
//                                the reference's example program (examples/mciExample.cxx): interpolate, median-smooth
//   mci_filter_test --copy in.nrrd out.nrrd   reader -> writer only (Image<short,3>, compressed); needs no GPU
//   mci_filter_test --api        interface checks that need no GPU (defaults, Modified(), error behaviour)
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "mci_b200_filter.hpp"
#include "mci_b200_io.hpp"

namespace
{

template <typename ImageType>
typename ImageType::Pointer
readRaw(const std::string & name, size_t x, size_t y, size_t z)
{
  typename ImageType::Pointer image = ImageType::New();
  image->SetRegions({ { x, y, z } });
  image->Allocate();
  std::ifstream f(name, std::ios::binary);
  if (!f)
    throw mci_b200::ExceptionObject(MCI_ERR_ARG, "cannot open " + name);
  f.read(reinterpret_cast<char *>(image->GetBufferPointer()),
         std::streamsize(image->GetNumberOfPixels() * sizeof(typename ImageType::PixelType)));
  if (size_t(f.gcount()) != image->GetNumberOfPixels() * sizeof(typename ImageType::PixelType))
    throw mci_b200::ExceptionObject(MCI_ERR_ARG, "short read from " + name);
  return image;
}

template <typename ImageType>
void
doTest(const std::string & inFilename, const std::string & outFilename, size_t x, size_t y, size_t z, bool UseDistanceTransform,
       bool ball, int axis, int label)
{
  typename ImageType::Pointer test = readRaw<ImageType>(inFilename, x, y, z);

  using mciType = mci_b200::MorphologicalContourInterpolator<ImageType>;
  typename mciType::Pointer mci = mciType::New();
  mci->SetInput(test);
  mci->SetUseDistanceTransform(UseDistanceTransform);
  mci->SetUseBallStructuringElement(ball);
  mci->SetAxis(axis);
  mci->SetLabel(typename ImageType::PixelType(label));
  mci->Update();

  // a second Update() without changes must not run again; a changed-and-restored parameter must give the same output
  const unsigned long launches = (unsigned long)mci->GetLastRunStatistics().n_launches;
  mci->Update();
  mci->SetAxis(axis == 0 ? 1 : 0);
  mci->SetAxis(axis);
  mci->Update();
  if (launches == 0 && mci->GetLastRunStatistics().n_segments > 0)
    throw mci_b200::ExceptionObject(MCI_ERR_STATE, "no kernel launches recorded");

  typename ImageType::Pointer out = mci->GetOutput();
  std::ofstream               f(outFilename, std::ios::binary);
  f.write(reinterpret_cast<const char *>(out->GetBufferPointer()),
          std::streamsize(out->GetNumberOfPixels() * sizeof(typename ImageType::PixelType)));
  if (!f)
    throw mci_b200::ExceptionObject(MCI_ERR_ARG, "cannot write " + outFilename);

  // slice indices found by the run: "axis label slice" lines on stdout
  auto indices = mci->GetLabeledSliceIndices();
  for (unsigned a = 0; a < indices.size(); ++a)
    for (const auto & kv : indices[a])
      for (auto s : kv.second)
        std::cout << a << " " << long(kv.first) << " " << long(s) << "\n";
}

// examples/mciExample.cxx:17-45 on raw files
template <typename ImageType>
void
doExample(const std::string & inFilename, const std::string & outFilename, size_t x, size_t y, size_t z, int smoothingRadius)
{
  typename ImageType::Pointer image = readRaw<ImageType>(inFilename, x, y, z);

  using mciType = mci_b200::MorphologicalContourInterpolator<ImageType>;
  typename mciType::Pointer mci = mciType::New();
  mci->SetInput(image);
  mci->Update();

  using MedianType = mci_b200::MedianImageFilter<ImageType, ImageType>;
  typename MedianType::Pointer medF = MedianType::New();
  medF->SetInput(mci->GetOutput());
  medF->SetRadius(smoothingRadius);
  medF->Update();

  typename ImageType::Pointer out = medF->GetOutput();
  std::ofstream               f(outFilename, std::ios::binary);
  f.write(reinterpret_cast<const char *>(out->GetBufferPointer()),
          std::streamsize(out->GetNumberOfPixels() * sizeof(typename ImageType::PixelType)));
  if (!f)
    throw mci_b200::ExceptionObject(MCI_ERR_ARG, "cannot write " + outFilename);
}

// test/itkMorphologicalContourInterpolationTest.cxx:25-49, on NRRD files
template <typename ImageType>
void
doFileTest(const std::string & inFilename, const std::string & outFilename, bool UseDistanceTransform, bool ball, int axis, int label)
{
  using ReaderType = mci_b200::ImageFileReader<ImageType>;
  typename ReaderType::Pointer reader = ReaderType::New();
  reader->SetFileName(inFilename);
  reader->Update();

  typename ImageType::Pointer test = reader->GetOutput();

  using mciType = mci_b200::MorphologicalContourInterpolator<ImageType>;
  typename mciType::Pointer mci = mciType::New();
  mci->SetInput(test);
  mci->SetUseDistanceTransform(UseDistanceTransform);
  mci->SetUseBallStructuringElement(ball);
  mci->SetAxis(axis);
  mci->SetLabel(typename ImageType::PixelType(label));
  mci->Update();

  using WriterType = mci_b200::ImageFileWriter<ImageType>;
  typename WriterType::Pointer writer = WriterType::New();
  writer->SetFileName(outFilename);
  writer->SetInput(mci->GetOutput());
  writer->SetUseCompression(true);
  writer->Update();
}

bool
endsWith(const std::string & s, const std::string & suffix)
{
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

#define EXPECT(cond)                                                                                                           \
  do                                                                                                                           \
  {                                                                                                                            \
    if (!(cond))                                                                                                               \
    {                                                                                                                          \
      std::cerr << "api check failed at line " << __LINE__ << ": " #cond << std::endl;                                         \
      return EXIT_FAILURE;                                                                                                     \
    }                                                                                                                          \
  } while (0)

int
apiChecks()
{
  using ImageType = mci_b200::Image<unsigned short, 3>;
  using mciType = mci_b200::MorphologicalContourInterpolator<ImageType>;
  mciType::Pointer mci = mciType::New();
  // defaults (H:230-235, X:84)
  EXPECT(mci->GetLabel() == 0);
  EXPECT(mci->GetAxis() == -1);
  EXPECT(mci->GetHeuristicAlignment() == true);
  EXPECT(mci->GetUseDistanceTransform() == true);
  EXPECT(mci->GetUseBallStructuringElement() == false);
  EXPECT(mci->GetUseCustomSlicePositions() == false);
  EXPECT(std::string(mci->GetNameOfClass()) == "MorphologicalContourInterpolator");
  EXPECT(mci->GetLabeledSliceIndices().size() == 3);
  // itkSetMacro: Modified() only when the value changes
  unsigned long t = mci->GetMTime();
  mci->SetAxis(-1);
  mci->SetUseBallStructuringElement(false);
  mci->SetUseExtrapolation(true);
  EXPECT(mci->GetMTime() == t);
  mci->SetAxis(2);
  EXPECT(mci->GetMTime() > t && mci->GetAxis() == 2);
  t = mci->GetMTime();
  mci->SetUseBallStructuringElement(true);
  EXPECT(mci->GetMTime() > t && mci->GetUseBallStructuringElement());
  // slice-index API (H:178-224)
  mci->SetLabeledSliceIndices(2, 5, std::vector<ImageType::IndexValueType>{ 9, 1, 4, 4 });
  EXPECT((mci->GetLabeledSliceIndices(2, 5) == mciType::SliceSetType{ 1, 4, 9 }));
  mci->SetLabeledSliceIndices(0, 7, mciType::SliceSetType{ 3, 8 });
  EXPECT(mci->GetLabeledSliceIndices()[0].at(7).size() == 2);
  t = mci->GetMTime();
  mci->ClearLabeledSliceIndices();
  EXPECT(mci->GetMTime() > t && mci->GetLabeledSliceIndices().size() == 3 && mci->GetLabeledSliceIndices()[2].empty());
  // Update() without input throws
  bool threw = false;
  try
  {
    mci->Update();
  }
  catch (mci_b200::ExceptionObject & error)
  {
    threw = error.GetCode() == MCI_ERR_STATE;
  }
  EXPECT(threw);
  // without a CUDA device Update() throws (no CPU fallback); with one it must succeed on an empty image
  ImageType::Pointer image = ImageType::New();
  image->SetRegions({ { 8, 8, 4 } });
  image->Allocate(true);
  mci->SetInput(image);
  bool ran = false;
  threw = false;
  try
  {
    mci->Update();
    ran = true;
  }
  catch (mci_b200::ExceptionObject & error)
  {
    threw = error.GetCode() == MCI_ERR_NO_DEVICE;
    std::cout << "no device: " << error.what() << std::endl;
  }
  EXPECT(ran != threw);
  if (ran)
  {
    EXPECT(mci->GetOutput()->GetNumberOfPixels() == 256);
    for (size_t k = 0; k < 256; ++k)
      EXPECT(mci->GetOutput()->GetBufferPointer()[k] == 0); // Empty input: output == input (test/CMakeLists.txt:91)
    std::cout << "device present: empty image passes through" << std::endl;
  }
  std::cout << "api ok" << std::endl;
  return EXIT_SUCCESS;
}

} // namespace

int
main(int argc, char * argv[])
{
  if (argc == 2 && std::string(argv[1]) == "--api")
    return apiChecks();
  if (argc >= 8 && std::string(argv[1]) == "--example")
  {
    const std::string type = argv[7];
    const size_t      x = std::strtoul(argv[4], nullptr, 10), y = std::strtoul(argv[5], nullptr, 10), z = std::strtoul(argv[6], nullptr, 10);
    const int         smoothingRadius = argc >= 9 ? std::stoi(argv[8]) : 2;
    try
    {
      if (type == "uchar")
        doExample<mci_b200::Image<unsigned char, 3>>(argv[2], argv[3], x, y, z, smoothingRadius);
      else if (type == "ushort")
        doExample<mci_b200::Image<unsigned short, 3>>(argv[2], argv[3], x, y, z, smoothingRadius);
      else
        doExample<mci_b200::Image<short, 3>>(argv[2], argv[3], x, y, z, smoothingRadius);
    }
    catch (mci_b200::ExceptionObject & exc)
    {
      std::cerr << exc.what() << std::endl;
      return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
  }
  if (argc == 4 && std::string(argv[1]) == "--copy")
  {
    // reader -> writer without the filter (needs no GPU): Image<short,3>, compression on
    try
    {
      auto reader = mci_b200::ImageFileReader<mci_b200::Image<short, 3>>::New();
      reader->SetFileName(argv[2]);
      reader->Update();
      auto writer = mci_b200::ImageFileWriter<mci_b200::Image<short, 3>>::New();
      writer->SetFileName(argv[3]);
      writer->SetInput(reader->GetOutput());
      writer->SetUseCompression(true);
      writer->Update();
    }
    catch (mci_b200::ExceptionObject & error)
    {
      std::cerr << "Error: " << error.what() << std::endl;
      return EXIT_FAILURE;
    }
    return EXIT_SUCCESS;
  }
  if (argc >= 3 && (endsWith(argv[1], ".nrrd") || mci_b200::detail::IsMetaImageName(argv[1])))
  {
    // the reference driver's command line (test/itkMorphologicalContourInterpolationTest.cxx:52-140)
    bool dt = false, ball = true;
    int  axis = -1, label = 0;
    if (argc >= 4)
    {
      char algo = char(std::toupper(argv[3][0]));
      if (algo == 'T')
        dt = true;
      else if (algo == 'C')
        ball = false;
    }
    if (argc >= 5)
      axis = int(std::strtol(argv[4], nullptr, 10));
    if (argc >= 6)
      label = int(std::strtol(argv[5], nullptr, 10));
    try
    {
      std::string pixelType;
      size_t      numDimensions = 0;
      mci_b200::ImageFileReader<mci_b200::Image<short, 3>>::ReadImageInformation(argv[1], pixelType, numDimensions);
      if (numDimensions == 3 && (pixelType == "short" || pixelType == "unsigned short"))
      {
        doFileTest<mci_b200::Image<short, 3>>(argv[1], argv[2], dt, ball, axis, label);
        return EXIT_SUCCESS;
      }
      std::cerr << "Unsupported image type:\n  Dimensions: " << numDimensions << "\n  Pixel type:" << pixelType << std::endl;
      return EXIT_FAILURE;
    }
    catch (mci_b200::ExceptionObject & error)
    {
      std::cerr << "Error: " << error.what() << std::endl;
      return EXIT_FAILURE;
    }
  }
  if (argc < 7)
  {
    std::cerr << "Usage: " << argv[0] << " inputRaw outputRaw X Y Z uchar|ushort|short [algorithm] [axis] [label]\n"
              << " algorithms:\n"
              << "  B = repeated dilations with ball structuring element\n"
              << "  C = repeated dilations with cross structuring element\n"
              << "  T = distance transform, ball (8-connected components)\n"
              << "  D = distance transform, cross (the filter's own defaults)\n"
              << " defaults: algo B, axis -1 (all axes), label 0 (all labels)" << std::endl;
    return EXIT_FAILURE;
  }
  const std::string in = argv[1], out = argv[2], type = argv[6];
  const size_t      x = std::strtoul(argv[3], nullptr, 10), y = std::strtoul(argv[4], nullptr, 10), z = std::strtoul(argv[5], nullptr, 10);
  bool              dt = false, ball = true; // as the reference driver: algo B unless told otherwise
  int               axis = -1, label = 0;
  if (argc >= 8)
  {
    char algo = char(std::toupper(argv[7][0]));
    if (algo == 'T')
      dt = true;
    else if (algo == 'C')
      ball = false;
    else if (algo == 'D')
    {
      dt = true;
      ball = false;
    }
  }
  if (argc >= 9)
    axis = int(std::strtol(argv[8], nullptr, 10));
  if (argc >= 10)
    label = int(std::strtol(argv[9], nullptr, 10));
  try
  {
    if (type == "uchar")
      doTest<mci_b200::Image<unsigned char, 3>>(in, out, x, y, z, dt, ball, axis, label);
    else if (type == "ushort")
      doTest<mci_b200::Image<unsigned short, 3>>(in, out, x, y, z, dt, ball, axis, label);
    else if (type == "short")
      doTest<mci_b200::Image<short, 3>>(in, out, x, y, z, dt, ball, axis, label);
    else
    {
      std::cerr << "Unsupported pixel type: " << type << std::endl;
      return EXIT_FAILURE;
    }
  }
  catch (mci_b200::ExceptionObject & error)
  {
    std::cerr << "Error: " << error.what() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
