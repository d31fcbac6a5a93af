// mci_b200_io.hpp -- NRRD (and MetaImage .mha / .mhd) file reader / writer for mci_b200::Image, shaped like itk::ImageFileReader /
// itk::ImageFileWriter as the reference's drivers use them (test/itkMorphologicalContourInterpolationTest.cxx:29-49,
// examples/mciExample.cxx:21-44).  Fixture I/O; needs zlib (-lz).
//
// The writer restates the byte layout of ITK's NrrdImageIO (teem's nrrdSave): header fields and their order, "%.17g"
// numbers, and for UseCompression teem's own gzip framing -- a 10-byte header without name or time stamp, raw deflate
// at zlib's default level with memLevel 8, CRC-32 and length trailer.  With it the files this library writes for the
// reference's regression inputs have the sha512 of the reference's regression baselines
// (test/Baseline/*.nrrd.sha512); tests/test_upstream_baselines.py checks exactly that.
#ifndef MCI_B200_IO_HPP
#define MCI_B200_IO_HPP

#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include <zlib.h>

#include "mci_b200_filter.hpp"

namespace mci_b200
{

namespace detail
{
template <typename TPixel>
struct NrrdTypeName;
template <>
struct NrrdTypeName<unsigned char>
{
  static const char *
  value()
  {
    return "unsigned char";
  }
};
template <>
struct NrrdTypeName<unsigned short>
{
  static const char *
  value()
  {
    return "unsigned short";
  }
};
template <>
struct NrrdTypeName<short>
{
  static const char *
  value()
  {
    return "short";
  }
};

struct NrrdHeader
{
  std::string                        type, encoding = "raw", endian = "little";
  int                                dimension = 0;
  std::vector<size_t>                sizes;
  std::vector<std::array<double, 3>> directions;
  std::array<double, 3>              origin{ { 0, 0, 0 } };
  size_t                             dataOffset = 0;
};

inline std::vector<double>
ParseVector(const std::string & text)
{
  std::vector<double> v;
  std::string         t = text;
  for (char & c : t)
    if (c == '(' || c == ')' || c == ',')
      c = ' ';
  std::istringstream in(t);
  double             d;
  while (in >> d)
    v.push_back(d);
  return v;
}

inline NrrdHeader
ParseNrrdHeader(const std::string & file, const std::string & name)
{
  if (file.compare(0, 4, "NRRD") != 0)
    throw ExceptionObject(MCI_ERR_ARG, name + ": not a NRRD file");
  const size_t cut = file.find("\n\n");
  if (cut == std::string::npos)
    throw ExceptionObject(MCI_ERR_ARG, name + ": NRRD header has no end");
  NrrdHeader         h;
  std::istringstream in(file.substr(0, cut));
  std::string        line;
  std::getline(in, line); // magic
  while (std::getline(in, line))
  {
    if (line.empty() || line[0] == '#')
      continue;
    const size_t colon = line.find(": ");
    if (colon == std::string::npos)
      continue;
    const std::string key = line.substr(0, colon), val = line.substr(colon + 2);
    if (key == "type")
      h.type = val;
    else if (key == "dimension")
      h.dimension = std::stoi(val);
    else if (key == "encoding")
      h.encoding = val;
    else if (key == "endian")
      h.endian = val;
    else if (key == "sizes")
      for (double d : ParseVector(val))
        h.sizes.push_back(size_t(d));
    else if (key == "space origin")
    {
      std::vector<double> v = ParseVector(val);
      for (size_t k = 0; k < v.size() && k < 3; ++k)
        h.origin[k] = v[k];
    }
    else if (key == "space directions")
    {
      std::vector<double> v = ParseVector(val);
      for (size_t k = 0; k + 2 < v.size(); k += 3)
        h.directions.push_back({ { v[k], v[k + 1], v[k + 2] } });
    }
  }
  h.dataOffset = cut + 2;
  return h;
}

inline std::string
Inflate(const char * data, size_t n, size_t expected, const std::string & name)
{
  std::string out(expected, '\0');
  z_stream    zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, 15 + 32) != Z_OK) // gzip or zlib framing
    throw ExceptionObject(MCI_ERR_STATE, "inflateInit2 failed");
  zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(data));
  zs.avail_in = uInt(n);
  zs.next_out = reinterpret_cast<Bytef *>(&out[0]);
  zs.avail_out = uInt(expected);
  const int rc = inflate(&zs, Z_FINISH);
  const size_t got = expected - zs.avail_out;
  inflateEnd(&zs);
  if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != expected)
    throw ExceptionObject(MCI_ERR_ARG, name + ": gzip payload does not hold the voxels the header announces");
  return out;
}

inline std::string
G17(double v)
{
  char buf[64];
  std::snprintf(buf, sizeof(buf), "%.17g", v);
  return buf;
}

// ---- MetaImage (.mha: header and voxels in one file; .mhd: header, voxels in the file ElementDataFile names) --------
// The container of the volumes the reference's Dice evaluation reads (doc/case_*_labels.csv, test/dscComparison.cxx).
// Covered: what ITK's MetaImageIO writes for a 3-D scalar label image -- MET_UCHAR / MET_USHORT / MET_SHORT, binary,
// either byte order, optional zlib compression.
inline bool
EndsWithNoCase(const std::string & s, const std::string & suffix)
{
  if (s.size() < suffix.size())
    return false;
  for (size_t k = 0; k < suffix.size(); ++k)
    if (std::tolower(static_cast<unsigned char>(s[s.size() - suffix.size() + k])) != suffix[k])
      return false;
  return true;
}
inline bool
IsMetaImageName(const std::string & name)
{
  return EndsWithNoCase(name, ".mha") || EndsWithNoCase(name, ".mhd");
}

struct MetaHeader
{
  int                   ndims = 0;
  std::vector<size_t>   sizes;
  std::string           elementType, dataFile;
  bool                  msb = false, compressed = false, binary = true;
  long long             compressedSize = -1, headerSize = 0;
  int                   channels = 1;
  std::array<double, 3> spacing{ { 1, 1, 1 } }, origin{ { 0, 0, 0 } };
  std::array<double, 9> matrix{ { 1, 0, 0, 0, 1, 0, 0, 0, 1 } };
  size_t                dataOffset = 0;
};

inline bool
MetaTrue(std::string v)
{
  for (char & c : v)
    c = char(std::tolower(static_cast<unsigned char>(c)));
  return v == "true" || v == "1" || v == "yes";
}

inline MetaHeader
ParseMetaHeader(const std::string & file, const std::string & name)
{
  MetaHeader h;
  size_t     pos = 0;
  bool       done = false;
  while (!done)
  {
    const size_t end = file.find('\n', pos);
    if (end == std::string::npos)
      throw ExceptionObject(MCI_ERR_ARG, name + ": MetaImage header without ElementDataFile");
    std::string line = file.substr(pos, end - pos);
    pos = end + 1;
    if (!line.empty() && line.back() == '\r')
      line.pop_back();
    const size_t eq = line.find('=');
    if (eq == std::string::npos)
      continue;
    auto trim = [](std::string t) {
      const size_t a = t.find_first_not_of(" \t"), b = t.find_last_not_of(" \t");
      return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
    };
    const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
    if (key == "ObjectType" && val != "Image")
      throw ExceptionObject(MCI_ERR_ARG, name + ": ObjectType " + val + " is not an image");
    else if (key == "NDims")
      h.ndims = std::stoi(val);
    else if (key == "DimSize")
      for (double d : ParseVector(val))
        h.sizes.push_back(size_t(d));
    else if (key == "ElementType")
      h.elementType = val;
    else if (key == "ElementNumberOfChannels")
      h.channels = std::stoi(val);
    else if (key == "BinaryData")
      h.binary = MetaTrue(val);
    else if (key == "BinaryDataByteOrderMSB" || key == "ElementByteOrderMSB")
      h.msb = MetaTrue(val);
    else if (key == "CompressedData")
      h.compressed = MetaTrue(val);
    else if (key == "CompressedDataSize")
      h.compressedSize = std::stoll(val);
    else if (key == "HeaderSize")
      h.headerSize = std::stoll(val);
    else if (key == "ElementSpacing" || key == "Offset" || key == "Position" || key == "Origin")
    {
      std::vector<double>     v = ParseVector(val);
      std::array<double, 3> & dst = key == "ElementSpacing" ? h.spacing : h.origin;
      for (size_t k = 0; k < v.size() && k < 3; ++k)
        dst[k] = v[k];
    }
    else if (key == "TransformMatrix" || key == "Rotation" || key == "Orientation")
    {
      std::vector<double> v = ParseVector(val);
      for (size_t k = 0; k < v.size() && k < 9; ++k)
        h.matrix[k] = v[k];
    }
    else if (key == "ElementDataFile") // always the last header line
    {
      h.dataFile = val;
      done = true;
    }
  }
  h.dataOffset = pos;
  return h;
}

template <typename TPixel>
struct MetaTypeName;
template <>
struct MetaTypeName<unsigned char>
{
  static const char *
  value()
  {
    return "MET_UCHAR";
  }
};
template <>
struct MetaTypeName<unsigned short>
{
  static const char *
  value()
  {
    return "MET_USHORT";
  }
};
template <>
struct MetaTypeName<short>
{
  static const char *
  value()
  {
    return "MET_SHORT";
  }
};

// numbers as the Python writer prints them (metaimage.py): integers without a fraction, anything else "%.17g"
inline std::string
MetaNumber(double v)
{
  if (v == double((long long)v))
    return std::to_string((long long)v);
  return G17(v);
}

inline std::string
DirName(const std::string & path)
{
  const size_t cut = path.find_last_of("/\\");
  return cut == std::string::npos ? std::string() : path.substr(0, cut + 1);
}
} // namespace detail

/** itk::ImageFileReader<TImage> for NRRD files (raw / gzip encodings, little endian, 3-D scalar). The voxels are
 *  converted to TImage::PixelType from the file's 8- or 16-bit integer type, as ITK's reader casts. */
template <typename TImage>
class ImageFileReader
{
public:
  using Pointer = std::shared_ptr<ImageFileReader>;
  static Pointer
  New()
  {
    return Pointer(new ImageFileReader());
  }
  void
  SetFileName(const std::string & n)
  {
    m_FileName = n;
  }
  /** Component type and dimension without reading the voxels (ImageIOBase::ReadImageInformation). */
  static void
  ReadImageInformation(const std::string & name, std::string & componentType, size_t & numDimensions)
  {
    std::ifstream f(name, std::ios::binary);
    if (!f)
      throw ExceptionObject(MCI_ERR_ARG, "cannot open " + name);
    std::string head(4096, '\0');
    f.read(&head[0], std::streamsize(head.size()));
    head.resize(size_t(f.gcount()));
    if (detail::IsMetaImageName(name))
    {
      detail::MetaHeader m = detail::ParseMetaHeader(head, name);
      componentType = m.elementType == "MET_UCHAR" ? "unsigned char"
                      : m.elementType == "MET_USHORT" ? "unsigned short"
                      : m.elementType == "MET_SHORT" ? "short"
                                                     : m.elementType;
      numDimensions = size_t(m.ndims);
      return;
    }
    detail::NrrdHeader h = detail::ParseNrrdHeader(head, name);
    componentType = h.type;
    numDimensions = size_t(h.dimension);
  }
  void
  Update()
  {
    std::ifstream f(m_FileName, std::ios::binary);
    if (!f)
      throw ExceptionObject(MCI_ERR_ARG, "cannot open " + m_FileName);
    std::string file((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    if (detail::IsMetaImageName(m_FileName))
    {
      ReadMetaImage(file);
      return;
    }
    detail::NrrdHeader h = detail::ParseNrrdHeader(file, m_FileName);
    if (h.dimension != 3 || h.sizes.size() != 3)
      throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": only 3-D volumes are supported");
    size_t bytesPer = 0;
    bool   isSigned = false;
    if (h.type == "unsigned char" || h.type == "uchar" || h.type == "uint8")
      bytesPer = 1;
    else if (h.type == "unsigned short" || h.type == "ushort" || h.type == "uint16")
      bytesPer = 2;
    else if (h.type == "short" || h.type == "signed short" || h.type == "int16")
    {
      bytesPer = 2;
      isSigned = true;
    }
    else
      throw ExceptionObject(MCI_ERR_PIXELTYPE, m_FileName + ": unsupported NRRD type " + h.type);
    if (bytesPer > 1 && h.endian != "little")
      throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": only little-endian files are supported");
    const size_t n = h.sizes[0] * h.sizes[1] * h.sizes[2];
    std::string  raw;
    if (h.encoding == "raw")
    {
      if (file.size() - h.dataOffset < n * bytesPer)
        throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": short read");
      raw.assign(file, h.dataOffset, n * bytesPer);
    }
    else if (h.encoding == "gzip" || h.encoding == "gz")
      raw = detail::Inflate(file.data() + h.dataOffset, file.size() - h.dataOffset, n * bytesPer, m_FileName);
    else
      throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": unsupported NRRD encoding " + h.encoding);
    m_Output = TImage::New();
    m_Output->SetRegions({ { h.sizes[0], h.sizes[1], h.sizes[2] } });
    m_Output->Allocate();
    m_Output->SetOrigin(h.origin);
    if (h.directions.size() == 3)
      m_Output->SetSpaceDirections({ { h.directions[0], h.directions[1], h.directions[2] } });
    typename TImage::PixelType * dst = m_Output->GetBufferPointer();
    for (size_t k = 0; k < n; ++k)
    {
      long v;
      if (bytesPer == 1)
        v = static_cast<unsigned char>(raw[k]);
      else
      {
        const unsigned u = static_cast<unsigned char>(raw[2 * k]) | (unsigned(static_cast<unsigned char>(raw[2 * k + 1])) << 8);
        v = isSigned ? long(int16_t(uint16_t(u))) : long(u);
      }
      dst[k] = static_cast<typename TImage::PixelType>(v);
    }
  }
  typename TImage::Pointer
  GetOutput()
  {
    return m_Output;
  }

private:
  ImageFileReader() = default;
  void
  ReadMetaImage(const std::string & file)
  {
    const detail::MetaHeader h = detail::ParseMetaHeader(file, m_FileName);
    if (h.ndims != 3 || h.sizes.size() != 3)
      throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": only 3-D volumes are supported");
    if (h.channels != 1 || !h.binary)
      throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": only binary scalar element data is supported");
    size_t bytesPer = 0;
    bool   isSigned = false;
    if (h.elementType == "MET_UCHAR")
      bytesPer = 1;
    else if (h.elementType == "MET_USHORT")
      bytesPer = 2;
    else if (h.elementType == "MET_SHORT")
    {
      bytesPer = 2;
      isSigned = true;
    }
    else
      throw ExceptionObject(MCI_ERR_PIXELTYPE, m_FileName + ": unsupported ElementType " + h.elementType);
    const size_t n = h.sizes[0] * h.sizes[1] * h.sizes[2];
    std::string  external;
    const char * data = file.data() + h.dataOffset;
    size_t       avail = file.size() - h.dataOffset;
    if (h.dataFile != "LOCAL")
    {
      const std::string path = detail::DirName(m_FileName) + h.dataFile;
      std::ifstream     g(path, std::ios::binary);
      if (!g)
        throw ExceptionObject(MCI_ERR_ARG, "cannot open " + path);
      external.assign((std::istreambuf_iterator<char>(g)), std::istreambuf_iterator<char>());
      const size_t skip = h.headerSize > 0 ? size_t(h.headerSize) : 0;
      if (skip > external.size())
        throw ExceptionObject(MCI_ERR_ARG, path + ": short read");
      data = external.data() + skip;
      avail = external.size() - skip;
    }
    std::string raw;
    if (h.compressed)
    {
      if (h.compressedSize >= 0 && size_t(h.compressedSize) < avail)
        avail = size_t(h.compressedSize);
      raw = detail::Inflate(data, avail, n * bytesPer, m_FileName);
    }
    else
    {
      if (avail < n * bytesPer)
        throw ExceptionObject(MCI_ERR_ARG, m_FileName + ": short read");
      raw.assign(data, n * bytesPer);
    }
    m_Output = TImage::New();
    m_Output->SetRegions({ { h.sizes[0], h.sizes[1], h.sizes[2] } });
    m_Output->Allocate();
    m_Output->SetOrigin(h.origin);
    std::array<std::array<double, 3>, 3> dirs;
    for (int a = 0; a < 3; ++a)
      for (int c = 0; c < 3; ++c)
        dirs[a][c] = h.matrix[3 * a + c] * h.spacing[a];
    m_Output->SetSpaceDirections(dirs);
    typename TImage::PixelType * dst = m_Output->GetBufferPointer();
    const int lo = h.msb ? 1 : 0, hi = h.msb ? 0 : 1;
    for (size_t k = 0; k < n; ++k)
    {
      long v;
      if (bytesPer == 1)
        v = static_cast<unsigned char>(raw[k]);
      else
      {
        const unsigned u = static_cast<unsigned char>(raw[2 * k + lo]) | (unsigned(static_cast<unsigned char>(raw[2 * k + hi])) << 8);
        v = isSigned ? long(int16_t(uint16_t(u))) : long(u);
      }
      dst[k] = static_cast<typename TImage::PixelType>(v);
    }
  }
  std::string             m_FileName;
  typename TImage::Pointer m_Output;
};

/** itk::ImageFileWriter<TImage> for NRRD files, byte-compatible with ITK's NrrdImageIO. */
template <typename TImage>
class ImageFileWriter
{
public:
  using Pointer = std::shared_ptr<ImageFileWriter>;
  static Pointer
  New()
  {
    return Pointer(new ImageFileWriter());
  }
  void
  SetFileName(const std::string & n)
  {
    m_FileName = n;
  }
  void
  SetInput(const typename TImage::Pointer & image)
  {
    m_Input = image;
  }
  void
  SetUseCompression(bool on)
  {
    m_UseCompression = on;
  }
  void
  Update()
  {
    using PixelType = typename TImage::PixelType;
    if (!m_Input)
      throw ExceptionObject(MCI_ERR_STATE, "ImageFileWriter: no input");
    const auto &       size = m_Input->GetSize();
    const auto &       dirs = m_Input->GetSpaceDirections();
    const auto &       org = m_Input->GetOrigin();
    if (detail::IsMetaImageName(m_FileName))
    {
      WriteMetaImage();
      return;
    }
    std::ostringstream head;
    auto vec = [](const std::array<double, 3> & v) { return "(" + detail::G17(v[0]) + "," + detail::G17(v[1]) + "," + detail::G17(v[2]) + ")"; };
    head << "NRRD0004\n# Complete NRRD file format specification at:\n# http://teem.sourceforge.net/nrrd/format.html\n";
    head << "type: " << detail::NrrdTypeName<PixelType>::value() << "\ndimension: 3\nspace: left-posterior-superior\n";
    head << "sizes: " << size[0] << " " << size[1] << " " << size[2] << "\n";
    head << "space directions: " << vec(dirs[0]) << " " << vec(dirs[1]) << " " << vec(dirs[2]) << "\n";
    head << "kinds: domain domain domain\n";
    if (sizeof(PixelType) > 1)
      head << "endian: little\n";
    head << "encoding: " << (m_UseCompression ? "gzip" : "raw") << "\n";
    head << "space origin: " << vec(org) << "\n\n";
    const size_t nbytes = m_Input->GetNumberOfPixels() * sizeof(PixelType); // host is little endian (x86-64)
    const char * raw = reinterpret_cast<const char *>(m_Input->GetBufferPointer());
    std::ofstream f(m_FileName, std::ios::binary);
    f << head.str();
    if (!m_UseCompression)
      f.write(raw, std::streamsize(nbytes));
    else
    {
      // teem's gzip framing (_nrrdGzOpen / _nrrdGzClose): magic, deflate, no flags, no time, no extra flags, unix
      const unsigned char gzHead[10] = { 0x1f, 0x8b, Z_DEFLATED, 0, 0, 0, 0, 0, 0, 0x03 };
      f.write(reinterpret_cast<const char *>(gzHead), 10);
      z_stream zs;
      std::memset(&zs, 0, sizeof(zs));
      if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -MAX_WBITS, 8, Z_DEFAULT_STRATEGY) != Z_OK)
        throw ExceptionObject(MCI_ERR_STATE, "deflateInit2 failed");
      std::vector<unsigned char> buf(1 << 16);
      zs.next_in = reinterpret_cast<Bytef *>(const_cast<char *>(raw));
      zs.avail_in = uInt(nbytes);
      int rc;
      do
      {
        zs.next_out = buf.data();
        zs.avail_out = uInt(buf.size());
        rc = deflate(&zs, Z_FINISH);
        f.write(reinterpret_cast<const char *>(buf.data()), std::streamsize(buf.size() - zs.avail_out));
      } while (rc == Z_OK || rc == Z_BUF_ERROR);
      deflateEnd(&zs);
      if (rc != Z_STREAM_END)
        throw ExceptionObject(MCI_ERR_STATE, "deflate failed");
      const uint32_t crc = uint32_t(crc32(crc32(0L, Z_NULL, 0), reinterpret_cast<const Bytef *>(raw), uInt(nbytes)));
      const uint32_t len = uint32_t(nbytes & 0xFFFFFFFFu);
      unsigned char  tail[8];
      for (int k = 0; k < 4; ++k)
      {
        tail[k] = (unsigned char)(crc >> (8 * k));
        tail[4 + k] = (unsigned char)(len >> (8 * k));
      }
      f.write(reinterpret_cast<const char *>(tail), 8);
    }
    if (!f)
      throw ExceptionObject(MCI_ERR_ARG, "cannot write " + m_FileName);
  }

private:
  ImageFileWriter() = default;
  // .mha: header and voxels in one file; .mhd: the voxels go to <stem>.raw / <stem>.zraw next to it.  Field order as
  // ITK's MetaImageIO writes it; zlib at MetaIO's level 2 when UseCompression is on.
  void
  WriteMetaImage()
  {
    using PixelType = typename TImage::PixelType;
    const auto & size = m_Input->GetSize();
    const auto & dirs = m_Input->GetSpaceDirections();
    const auto & org = m_Input->GetOrigin();
    const size_t nbytes = m_Input->GetNumberOfPixels() * sizeof(PixelType); // host is little endian (x86-64)
    const char * raw = reinterpret_cast<const char *>(m_Input->GetBufferPointer());
    std::vector<unsigned char> packed;
    if (m_UseCompression)
    {
      uLongf cap = compressBound(uLong(nbytes));
      packed.resize(cap);
      if (compress2(packed.data(), &cap, reinterpret_cast<const Bytef *>(raw), uLong(nbytes), 2) != Z_OK)
        throw ExceptionObject(MCI_ERR_STATE, "compress2 failed");
      packed.resize(cap);
    }
    const bool  separate = detail::EndsWithNoCase(m_FileName, ".mhd");
    std::string dataName = "LOCAL";
    if (separate)
    {
      const size_t slash = m_FileName.find_last_of("/\\");
      std::string  stem = slash == std::string::npos ? m_FileName : m_FileName.substr(slash + 1);
      stem.resize(stem.size() - 4);
      dataName = stem + (m_UseCompression ? ".zraw" : ".raw");
    }
    double spacing[3];
    double matrix[9];
    for (int a = 0; a < 3; ++a)
    {
      spacing[a] = std::sqrt(dirs[a][0] * dirs[a][0] + dirs[a][1] * dirs[a][1] + dirs[a][2] * dirs[a][2]);
      if (spacing[a] == 0)
        spacing[a] = 1;
      for (int c = 0; c < 3; ++c)
        matrix[3 * a + c] = dirs[a][c] / spacing[a];
    }
    std::ostringstream head;
    head << "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\n";
    head << "CompressedData = " << (m_UseCompression ? "True" : "False") << "\n";
    if (m_UseCompression)
      head << "CompressedDataSize = " << packed.size() << "\n";
    head << "TransformMatrix =";
    for (int k = 0; k < 9; ++k)
      head << " " << detail::MetaNumber(matrix[k]);
    head << "\nOffset = " << detail::MetaNumber(org[0]) << " " << detail::MetaNumber(org[1]) << " " << detail::MetaNumber(org[2]);
    head << "\nCenterOfRotation = 0 0 0\nAnatomicalOrientation = RAI\n";
    head << "ElementSpacing = " << detail::MetaNumber(spacing[0]) << " " << detail::MetaNumber(spacing[1]) << " "
         << detail::MetaNumber(spacing[2]) << "\n";
    head << "DimSize = " << size[0] << " " << size[1] << " " << size[2] << "\n";
    head << "ElementType = " << detail::MetaTypeName<PixelType>::value() << "\nElementDataFile = " << dataName << "\n";
    std::ofstream f(m_FileName, std::ios::binary);
    f << head.str();
    std::ofstream   g;
    std::ofstream & body = separate ? g : f;
    if (separate)
      g.open(detail::DirName(m_FileName) + dataName, std::ios::binary);
    if (m_UseCompression)
      body.write(reinterpret_cast<const char *>(packed.data()), std::streamsize(packed.size()));
    else
      body.write(raw, std::streamsize(nbytes));
    if (!f || !body)
      throw ExceptionObject(MCI_ERR_ARG, "cannot write " + m_FileName);
  }
  std::string             m_FileName;
  typename TImage::Pointer m_Input;
  bool                    m_UseCompression{ false };
};

} // namespace mci_b200
#endif // MCI_B200_IO_HPP
