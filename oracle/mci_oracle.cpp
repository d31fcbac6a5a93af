// mci_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, NOT PRODUCT CODE).
//
// A single-file CPU restatement of itk::MorphologicalContourInterpolator's
// algorithm, following /root/reference/include/itkMorphologicalContourInterpolator.hxx
// ("X" below) function by function, with the ITK 5.4 primitives it composes
// (connected components, signed Maurer distance map, binary dilation) restated
// from their documented behaviour.
//
// PARITY: ITK is not installed in this environment, so the real reference cannot
// be built or run here (no oracle/_ref).  The oracle is pinned instead by
// (i) four of the reference's own regression baselines: SevenLabels_B,
// ManyToMany_B, ManyToMany_C and ManyToMany_T (test/Baseline/*.nrrd.sha512) --
// its output, written through ITK's NRRD byte layout, reproduces their sha512
// content links, i.e. is voxel-exact against output of the real ITK filter
// (tests/test_upstream_baselines.py; SevenLabels_C / _T are NOT reproduced, see
// DESIGN.md section 0); (ii) brute-force / scipy / independent restatements of
// each primitive (tests/test_oracle_primitives.py); (iii) the invariants the
// reference's own tests state (tests/test_oracle_invariants.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (libmci_b200.so) never
// links or calls it.
//
// Build: see oracle/Makefile  (g++ -O3 -march=native -std=c++17 -shared -fPIC -pthread)

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <set>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>
#include <immintrin.h>

namespace
{

typedef long IV; // itk::IndexValueType
typedef unsigned long IdT; // itk::IdentifierType

// ---- minimal 2-D region / image (stand-ins for itk::ImageRegion<2>, itk::Image<P,2>) ----
struct Rect
{
  IV idx[2];
  IV sz[2];
  IV npix() const { return sz[0] * sz[1]; }
  bool inside(IV x, IV y) const
  {
    return x >= idx[0] && x < idx[0] + sz[0] && y >= idx[1] && y < idx[1] + sz[1];
  }
};

template <typename P>
struct Slice
{
  Rect r;
  std::vector<P> px;
  void alloc(const Rect & rr)
  {
    r = rr;
    px.assign(size_t(std::max<IV>(rr.sz[0], 0)) * size_t(std::max<IV>(rr.sz[1], 0)), P(0));
  }
  P & at(IV x, IV y) { return px[size_t(y - r.idx[1]) * size_t(r.sz[0]) + size_t(x - r.idx[0])]; }
  const P & at(IV x, IV y) const { return px[size_t(y - r.idx[1]) * size_t(r.sz[0]) + size_t(x - r.idx[0])]; }
};
typedef Slice<uint8_t> BoolSlice;
typedef std::shared_ptr<BoolSlice> BoolPtr;

struct Box3
{
  IV idx[3];
  IV sz[3];
};

// X:107-126  ExpandRegion (rect grows to include a point)
static void
expand_rect(Rect & r, const IV p[2])
{
  for (int a = 0; a < 2; ++a)
  {
    if (r.idx[a] > p[a])
    {
      r.sz[a] = r.sz[a] + r.idx[a] - p[a];
      r.idx[a] = p[a];
    }
    else if (r.idx[a] + r.sz[a] <= p[a])
    {
      r.sz[a] = p[a] - r.idx[a] + 1;
    }
  }
}
static void
expand_box(Box3 & r, const IV p[3])
{
  for (int a = 0; a < 3; ++a)
  {
    if (r.idx[a] > p[a])
    {
      r.sz[a] = r.sz[a] + r.idx[a] - p[a];
      r.idx[a] = p[a];
    }
    else if (r.idx[a] + r.sz[a] <= p[a])
    {
      r.sz[a] = p[a] - r.idx[a] + 1;
    }
  }
}

// X:1011-1030 IntersectionRegions: overlap of (iRegion shifted by t) with jRegion, both modified.
static void
intersection_regions(const IV t[2], Rect & ri, Rect & rj)
{
  for (int d = 0; d < 2; ++d)
  {
    IV iSize = ri.sz[d], jSize = rj.sz[d];
    IV iBegin = ri.idx[d] + t[d];
    IV jBegin = rj.idx[d];
    IV m = std::max(iBegin, jBegin);
    ri.sz[d] = std::min(iSize - (m - iBegin), jSize - (m - jBegin));
    ri.idx[d] = m - t[d];
    rj.idx[d] = m;
  }
  rj.sz[0] = ri.sz[0];
  rj.sz[1] = ri.sz[1];
}

// ---- ITK primitive restatements ----

// A.1  ConnectedComponentImageFilter<bool 2-D -> P>, background 0: ids 1..n numbered by the raster
// position (slow index, then fast index) of each component's first pixel.
template <typename P>
static IdT
connected_components(const BoolSlice & m, bool fully, Slice<P> & out)
{
  const IV w = m.r.sz[0], h = m.r.sz[1];
  out.alloc(m.r);
  std::vector<int32_t> par(size_t(w) * size_t(h), -1);
  auto find = [&](int32_t a) {
    while (par[a] != a)
    {
      par[a] = par[par[a]];
      a = par[a];
    }
    return a;
  };
  auto link = [&](int32_t a, int32_t b) {
    a = find(a);
    b = find(b);
    if (a < b)
      par[b] = a;
    else if (b < a)
      par[a] = b;
  };
  for (IV y = 0; y < h; ++y)
    for (IV x = 0; x < w; ++x)
    {
      int32_t p = int32_t(y * w + x);
      if (!m.px[p])
        continue;
      par[p] = p;
      if (x > 0 && m.px[p - 1])
        link(p, p - 1);
      if (y > 0)
      {
        if (m.px[p - w])
          link(p, int32_t(p - w));
        if (fully)
        {
          if (x > 0 && m.px[p - w - 1])
            link(p, int32_t(p - w - 1));
          if (x + 1 < w && m.px[p - w + 1])
            link(p, int32_t(p - w + 1));
        }
      }
    }
  std::vector<int32_t> ids(size_t(w) * size_t(h), 0);
  IdT n = 0;
  for (size_t p = 0; p < par.size(); ++p)
  {
    if (par[p] < 0)
      continue;
    int32_t r = find(int32_t(p));
    if (size_t(r) == p)
      ids[p] = int32_t(++n); // roots are met in raster order before any of their members
    out.px[p] = P(ids[r]);
  }
  if (n > IdT(std::numeric_limits<P>::max()))
    throw std::runtime_error("connected components: object count exceeds pixel type");
  return n;
}

static const int64_t EDT_INF = int64_t(1) << 40;

// A.2  exact squared Euclidean distance to the nearest set pixel (what SignedMaurerDistanceMap with
// SquaredDistance=true, UseImageSpacing=false yields OUTSIDE the object).  out[p] = d^2, EDT_INF if the
// mask is empty.  Separable: column pass, then an exact lower envelope of parabolas per row.
static void
edt_squared(const BoolSlice & m, std::vector<int64_t> & d2)
{
  const IV w = m.r.sz[0], h = m.r.sz[1];
  const int64_t BIG = int64_t(1) << 20; // larger than any image side
  std::vector<int64_t> g(size_t(w) * size_t(h));
  for (IV x = 0; x < w; ++x)
  {
    int64_t d = BIG;
    for (IV y = 0; y < h; ++y)
    {
      d = m.px[y * w + x] ? 0 : std::min(d + 1, BIG);
      g[y * w + x] = d;
    }
    d = BIG;
    for (IV y = h - 1; y >= 0; --y)
    {
      d = m.px[y * w + x] ? 0 : std::min(d + 1, BIG);
      g[y * w + x] = std::min(g[y * w + x], d);
    }
  }
  d2.assign(size_t(w) * size_t(h), EDT_INF);
  std::vector<IV> v(w);
  std::vector<double> z(w + 1);
  std::vector<int64_t> f(w);
  for (IV y = 0; y < h; ++y)
  {
    // only columns that contain a set pixel take part in the envelope
    IV n = 0;
    for (IV x = 0; x < w; ++x)
    {
      int64_t gy = g[y * w + x];
      if (gy >= BIG)
        continue;
      int64_t fx = gy * gy;
      // insert parabola (x, fx)
      while (n > 0)
      {
        IV q = v[n - 1];
        // intersection abscissa of parabolas q and x
        double s = double((fx + x * x) - (f[n - 1] + q * q)) / double(2 * (x - q));
        if (s <= z[n - 1])
          --n;
        else
        {
          z[n] = s;
          break;
        }
      }
      if (n == 0)
        z[0] = -1e30;
      v[n] = x;
      f[n] = fx;
      ++n;
    }
    if (n == 0)
      continue;
    IV k = 0;
    for (IV x = 0; x < w; ++x)
    {
      while (k + 1 < n && z[k + 1] < double(x))
        ++k;
      // exactness: take the min over the envelope neighbour candidates as well (guards the
      // floating-point breakpoints; the minimum over parabolas is what is defined)
      int64_t best = (x - v[k]) * (x - v[k]) + f[k];
      if (k + 1 < n)
        best = std::min(best, (x - v[k + 1]) * (x - v[k + 1]) + f[k + 1]);
      if (k > 0)
        best = std::min(best, (x - v[k - 1]) * (x - v[k - 1]) + f[k - 1]);
      d2[y * w + x] = best;
    }
  }
}

// Brute-force O(A^2) version, used by the tests to pin edt_squared.
static void
edt_squared_brute(const BoolSlice & m, std::vector<int64_t> & d2)
{
  const IV w = m.r.sz[0], h = m.r.sz[1];
  std::vector<std::pair<IV, IV>> on;
  for (IV y = 0; y < h; ++y)
    for (IV x = 0; x < w; ++x)
      if (m.px[y * w + x])
        on.push_back({ x, y });
  d2.assign(size_t(w) * size_t(h), EDT_INF);
  for (IV y = 0; y < h; ++y)
    for (IV x = 0; x < w; ++x)
    {
      int64_t b = EDT_INF;
      for (auto & q : on)
        b = std::min<int64_t>(b, (q.first - x) * (q.first - x) + (q.second - y) * (q.second - y));
      d2[y * w + x] = b;
    }
}

// A.3  BinaryDilateImageFilter with a radius-1 cross (centre + 4 face neighbours) or radius-1
// "ball" (= full 3x3 box); pixels outside the region are background.
static BoolPtr
dilate_r1(const BoolSlice & s, bool ball)
{
  BoolPtr o = std::make_shared<BoolSlice>();
  o->alloc(s.r);
  const IV w = s.r.sz[0], h = s.r.sz[1];
  for (IV y = 0; y < h; ++y)
    for (IV x = 0; x < w; ++x)
    {
      bool v = s.px[y * w + x];
      if (!v)
      {
        for (IV dy = -1; dy <= 1 && !v; ++dy)
          for (IV dx = -1; dx <= 1 && !v; ++dx)
          {
            if (!ball && dx != 0 && dy != 0)
              continue;
            IV xx = x + dx, yy = y + dy;
            if (xx < 0 || yy < 0 || xx >= w || yy >= h)
              continue;
            v = s.px[yy * w + xx];
          }
      }
      o->px[y * w + x] = v;
    }
  return o;
}

// A.5  `PixelType dist = fractioning * sdf` (X:426,431): float multiply then float->integer conversion.
// Out of range is UB in C++; g++ on x86-64 emits cvttss2si (int32, 0x80000000 when out of range) and
// keeps the low bits.  Stated explicitly here so every compiler agrees with what the reference build does.
template <typename P>
static inline P
dist_bin(float sdf)
{
  const short fractioning = 10;
  float prod = fractioning * sdf;
  int32_t q = _mm_cvttss_si32(_mm_set_ss(prod));
  return P(uint32_t(q)); // modular narrowing
}

struct Params
{
  int64_t label; // 0 = all
  int32_t axis; // -1 = all
  int32_t heuristic_alignment;
  int32_t use_distance_transform;
  int32_t use_ball;
  int32_t use_extrapolation;
  int32_t use_custom_slice_positions;
  int32_t threads;
  int32_t shard_index, shard_count; // interpolate only segments k with k % shard_count == shard_index (0/1 = all)
};

struct Stats
{
  int64_t segments;
  int64_t root_jobs; // non-recursive Interpolate1to1 calls
  int64_t nodes; // all Interpolate1to1 calls
  int64_t align_calls;
  int64_t align_evals;
  int64_t extrapolations;
  int64_t one_to_n;
  double t_detect, t_cc, t_interp, t_final, t_total;
};

template <typename T>
class Oracle
{
public:
  typedef Slice<T> ConnSlice;
  typedef std::shared_ptr<ConnSlice> ConnPtr;
  typedef std::vector<T> PixelList;

  const T * in = nullptr;
  T * out = nullptr;
  IV dims[3] = { 0, 0, 0 };
  Params P{};
  std::vector<std::map<T, std::set<IV>>> labeledSlices;
  std::map<T, Box3> bboxes;
  IdT minAlignIters = 8; // pow(2, ImageDimension=3)   X:87
  IdT maxAlignIters = 216; // pow(6, 3)                 X:89
  std::mutex writeMutex; // X:743-753
  std::atomic<int64_t> nRoot{ 0 }, nNodes{ 0 }, nAlign{ 0 }, nEval{ 0 }, nExtra{ 0 }, n1N{ 0 };
  FILE * trace = nullptr;
  std::mutex traceMutex;

  T vox(IV x, IV y, IV z) const { return in[(size_t(z) * size_t(dims[1]) + size_t(y)) * size_t(dims[0]) + size_t(x)]; }

  // X:128-201
  void DetermineSliceOrientations()
  {
    labeledSlices.clear();
    labeledSlices.resize(3);
    bboxes.clear();
    for (IV z = 0; z < dims[2]; ++z)
      for (IV y = 0; y < dims[1]; ++y)
        for (IV x = 0; x < dims[0]; ++x)
        {
          const T val = vox(x, y, z);
          if (val == 0)
            continue;
          const IV ind[3] = { x, y, z };
          Box3 one;
          for (int a = 0; a < 3; ++a)
          {
            one.idx[a] = ind[a];
            one.sz[a] = 1;
          }
          auto res = bboxes.insert(std::make_pair(val, one));
          if (!res.second)
            expand_box(res.first->second, ind);

          unsigned cTrue = 0, cAdjacent = 0, axis = 0;
          for (unsigned a = 0; a < 3; ++a)
          {
            IV pi[3] = { x, y, z }, ni[3] = { x, y, z };
            pi[a]--;
            ni[a]++;
            T prev = 0, next = 0;
            if (pi[a] >= 0)
              prev = vox(pi[0], pi[1], pi[2]);
            if (ni[a] < dims[a])
              next = vox(ni[0], ni[1], ni[2]);
            if (prev == 0 && next == 0)
            {
              axis = a;
              ++cTrue;
            }
            else if (prev == val && next == val)
              ++cAdjacent;
          }
          if (cTrue == 1 && cAdjacent == 2)
            if (P.axis == -1 || P.axis == int(axis))
              labeledSlices[axis][val].insert(ind[axis]);
        }
  }

  // X:1224-1238: Extract (axis collapsed) -> == label -> connected components
  ConnPtr RegionedConnectedComponents(const Box3 & region, int axis, T label, IdT & count)
  {
    int d0 = axis == 0 ? 1 : 0;
    int d1 = axis == 2 ? 1 : 2;
    BoolSlice m;
    Rect r;
    r.idx[0] = region.idx[d0];
    r.idx[1] = region.idx[d1];
    r.sz[0] = region.sz[d0];
    r.sz[1] = region.sz[d1];
    m.alloc(r);
    IV p[3];
    p[axis] = region.idx[axis];
    for (IV v = 0; v < r.sz[1]; ++v)
      for (IV u = 0; u < r.sz[0]; ++u)
      {
        p[d0] = r.idx[0] + u;
        p[d1] = r.idx[1] + v;
        m.px[v * r.sz[0] + u] = (vox(p[0], p[1], p[2]) == label);
      }
    ConnPtr c = std::make_shared<ConnSlice>();
    count = connected_components<T>(m, P.use_ball != 0, *c);
    return c;
  }

  // X:1101-1134
  void Centroid(const ConnSlice & conn, const PixelList & ids, IV ret[2])
  {
    IV ind[2] = { 0, 0 };
    IdT pixelCount = 0;
    const IV w = conn.r.sz[0], h = conn.r.sz[1];
    for (IV y = 0; y < h; ++y)
      for (IV x = 0; x < w; ++x)
      {
        T val = conn.px[y * w + x];
        if (val && std::find(ids.begin(), ids.end(), val) != ids.end())
        {
          ++pixelCount;
          ind[0] += conn.r.idx[0] + x;
          ind[1] += conn.r.idx[1] + y;
        }
      }
    for (int d = 0; d < 2; ++d)
      ret[d] = IV(IdT(ind[d]) / pixelCount); // signed / unsigned, as in the reference expression
  }

  // X:1032-1077
  IdT Intersection(const ConnSlice & iConn, T iId, const ConnSlice & jConn, const PixelList & jIds, const IV t[2])
  {
    Rect ri = iConn.r, rj = jConn.r;
    intersection_regions(t, ri, rj);
    std::vector<IdT> counts(jIds.size(), 0);
    for (IV y = 0; y < ri.sz[1]; ++y)
      for (IV x = 0; x < ri.sz[0]; ++x)
      {
        if (iConn.at(ri.idx[0] + x, ri.idx[1] + y) == iId)
        {
          T jVal = jConn.at(rj.idx[0] + x, rj.idx[1] + y);
          auto res = std::find(jIds.begin(), jIds.end(), jVal);
          if (res != jIds.end())
            ++counts[res - jIds.begin()];
        }
      }
    IdT sum = 0;
    for (size_t k = 0; k < jIds.size(); ++k)
    {
      if (counts[k] == 0)
        return 0;
      sum += counts[k];
    }
    return sum;
  }

  // X:1136-1222
  void Align(const ConnSlice & iConn, T iId, const ConnSlice & jConn, const PixelList & jIds, IV best[2])
  {
    ++nAlign;
    PixelList iIds(1, iId);
    IV ic[2], jc[2], ind[2];
    Centroid(iConn, iIds, ic);
    Centroid(jConn, jIds, jc);
    for (int d = 0; d < 2; ++d)
      ind[d] = jc[d] - ic[d];

    Rect sr;
    for (int d = 0; d < 2; ++d)
    {
      sr.idx[d] = jConn.r.idx[d] - iConn.r.idx[d] - iConn.r.sz[d] + 1;
      sr.sz[d] = iConn.r.sz[d] + jConn.r.sz[d] - 1;
    }
    BoolSlice searched;
    searched.alloc(sr);

    std::queue<std::pair<IV, IV>> q;
    q.push({ 0, 0 });
    q.push({ ind[0], ind[1] });
    searched.at(0, 0) = 1;
    searched.at(ind[0], ind[1]) = 1;
    IdT score, maxScore = 0, iter = 0;
    best[0] = best[1] = 0; // (uninitialised in the reference; unreachable for non-empty shapes, X:1180)
    IdT minIter = std::min<IdT>(minAlignIters, IdT(sr.npix()));
    IdT maxIter = std::max<IdT>(maxAlignIters, IdT(std::sqrt(double(sr.npix()))));
    while (!q.empty())
    {
      ind[0] = q.front().first;
      ind[1] = q.front().second;
      q.pop();
      score = Intersection(iConn, iId, jConn, jIds, ind);
      ++nEval;
      if (score > maxScore)
      {
        maxScore = score;
        best[0] = ind[0];
        best[1] = ind[1];
      }
      if (!P.heuristic_alignment || maxScore == 0 || iter <= minIter || (score > maxScore * 0.9 && iter <= maxIter))
      {
        for (int d = 0; d < 2; ++d)
        {
          ind[d] -= 1;
          if (sr.inside(ind[0], ind[1]) && !searched.at(ind[0], ind[1]))
          {
            q.push({ ind[0], ind[1] });
            searched.at(ind[0], ind[1]) = 1;
            ++iter;
          }
          ind[d] += 2;
          if (sr.inside(ind[0], ind[1]) && !searched.at(ind[0], ind[1]))
          {
            q.push({ ind[0], ind[1] });
            searched.at(ind[0], ind[1]) = 1;
            ++iter;
          }
          ind[d] -= 1;
        }
      }
    }
  }

  // X:994-1009
  ConnPtr TranslateImage(const ConnSlice & image, const IV t[2], Rect newRegion)
  {
    ConnPtr res = std::make_shared<ConnSlice>();
    res->alloc(newRegion);
    Rect inR = image.r;
    intersection_regions(t, inR, newRegion);
    for (IV y = 0; y < inR.sz[1]; ++y)
      for (IV x = 0; x < inR.sz[0]; ++x)
        res->at(newRegion.idx[0] + x, newRegion.idx[1] + y) = image.at(inR.idx[0] + x, inR.idx[1] + y);
    return res;
  }

  // X:518-554
  Rect BoundingBox(const ConnSlice & image)
  {
    Rect nr = image.r;
    IV mn[2] = { nr.idx[0] + nr.sz[0], nr.idx[1] + nr.sz[1] };
    IV mx[2] = { nr.idx[0], nr.idx[1] };
    for (IV y = 0; y < nr.sz[1]; ++y)
      for (IV x = 0; x < nr.sz[0]; ++x)
        if (image.px[y * nr.sz[0] + x])
        {
          IV p[2] = { nr.idx[0] + x, nr.idx[1] + y };
          for (int d = 0; d < 2; ++d)
          {
            mn[d] = std::min(mn[d], p[d]);
            mx[d] = std::max(mx[d], p[d]);
          }
        }
    for (int d = 0; d < 2; ++d)
    {
      nr.idx[d] = mn[d];
      nr.sz[d] = mx[d] - mn[d] + 1;
    }
    return nr;
  }

  // X:253-320
  BoolPtr Dilate1(const BoolPtr & seed, const BoolPtr & mask)
  {
    BoolPtr t = dilate_r1(*seed, P.use_ball != 0);
    for (size_t p = 0; p < t->px.size(); ++p)
      t->px[p] = t->px[p] && mask->px[p];
    return t;
  }

  // X:322-338
  std::vector<BoolPtr> GenerateDilationSequence(const BoolPtr & begin, const BoolPtr & end)
  {
    std::vector<BoolPtr> seq;
    seq.push_back(Dilate1(begin, end));
    do
    {
      seq.push_back(Dilate1(seq.back(), end));
    } while (seq.back()->px != seq[seq.size() - 2]->px); // ImagesEqual X:54-80
    seq.pop_back();
    return seq;
  }

  // X:1079-1099
  static IdT CardinalSymmetricDifference(const BoolSlice & a, const BoolSlice & b)
  {
    IdT c = 0;
    for (size_t p = 0; p < a.px.size(); ++p)
      c += (a.px[p] != b.px[p]);
    return c;
  }

  // X:340-388
  BoolPtr FindMedianImageDilations(const BoolPtr & intersection, const BoolPtr & iMask, const BoolPtr & jMask)
  {
    std::vector<BoolPtr> iSeq = GenerateDilationSequence(intersection, iMask);
    std::vector<BoolPtr> jSeq = GenerateDilationSequence(intersection, jMask);
    std::reverse(iSeq.begin(), iSeq.end());
    if (iSeq.size() < jSeq.size())
      iSeq.swap(jSeq);
    float ratio = float(jSeq.size()) / iSeq.size();

    std::vector<BoolPtr> seq;
    for (unsigned x = 0; x < iSeq.size(); x++)
    {
      unsigned xj = ratio * x;
      BoolPtr o = std::make_shared<BoolSlice>();
      o->alloc(iMask->r);
      for (size_t p = 0; p < o->px.size(); ++p)
        o->px[p] = iSeq[x]->px[p] || jSeq[xj]->px[p];
      seq.push_back(o);
    }
    unsigned minIndex = 0;
    IdT mn = IdT(iMask->r.npix());
    for (unsigned x = 0; x < iSeq.size(); x++)
    {
      IdT iS = CardinalSymmetricDifference(*seq[x], *iMask);
      IdT jS = CardinalSymmetricDifference(*seq[x], *jMask);
      IdT xScore = iS >= jS ? iS - jS : jS - iS;
      if (xScore < mn)
      {
        mn = xScore;
        minIndex = x;
      }
    }
    return seq[minIndex];
  }

  // X:390-516 (MaurerDM + histograms + threshold)
  BoolPtr FindMedianImageDistances(const BoolPtr & intersection, const BoolPtr & iMask, const BoolPtr & jMask)
  {
    std::vector<int64_t> d2;
    edt_squared(*intersection, d2);
    const size_t N = iMask->px.size();
    // sdf as the float image the reference sees.  SignedMaurerDistanceMapImageFilter is used with its default
    // SquaredDistance = false (X:395-402 never sets it), so outside pixels hold the Euclidean distance itself:
    // ITK's last pass writes float(std::sqrt(double(|d^2|))).  Inside pixels are <= 0 (only the sign matters
    // here), FLT_MAX if the object is empty.  Pinned by the upstream baseline ManyToMany_T.nrrd (whose sha512
    // is reproduced with this convention and with no other; tests/test_upstream_baselines.py).
    std::vector<float> sdf(N);
    for (size_t p = 0; p < N; ++p)
      sdf[p] = intersection->px[p]
                 ? -1.0f
                 : (d2[p] >= EDT_INF ? std::numeric_limits<float>::max() : float(std::sqrt(double(float(d2[p])))));

    BoolPtr orImage = std::make_shared<BoolSlice>();
    orImage->alloc(iMask->r);
    std::vector<long long> iHist, jHist;
    for (size_t p = 0; p < N; ++p)
    {
      bool iM = iMask->px[p], jM = jMask->px[p];
      if (iM != jM)
      {
        T dist = dist_bin<T>(sdf[p]);
        if (std::numeric_limits<T>::is_signed && dist < 0)
          throw std::runtime_error("distance bin negative for signed pixel type (reference UB, X:431-437)");
        std::vector<long long> & H = iM ? iHist : jHist;
        if (size_t(dist) >= H.size())
          H.resize(size_t(dist) + 1, 0);
        H[size_t(dist)]++;
        orImage->px[p] = 1;
      }
      else if (iM && jM)
        orImage->px[p] = 1;
    }
    size_t maxSize = std::max(iHist.size(), jHist.size());
    if (maxSize == 0)
      return intersection;
    iHist.resize(maxSize, 0);
    jHist.resize(maxSize, 0);
    std::vector<long long> iSum(maxSize, 0), jSum(maxSize, 0);
    iSum[0] = iHist[0];
    jSum[0] = jHist[0];
    for (size_t b = 1; b < maxSize; b++)
    {
      iSum[b] = iSum[b - 1] + iHist[b];
      jSum[b] = jSum[b - 1] + jHist[b];
    }
    long long iTotal = iSum[maxSize - 1], jTotal = jSum[maxSize - 1];
    int bestBin = 0;
    long long bestDiff = LLONG_MAX;
    for (size_t b = 0; b < maxSize; b++)
    {
      long long iS = std::llabs(iTotal - iSum[b] + jSum[b]);
      long long jS = std::llabs(jTotal - jSum[b] + iSum[b]);
      long long diff = std::llabs(iS - jS);
      if (diff < bestDiff)
      {
        bestDiff = diff;
        bestBin = int(b);
      }
    }
    const short fractioning = 10;
    float upper = float(bestBin) / fractioning;
    BoolPtr median = std::make_shared<BoolSlice>();
    median->alloc(iMask->r);
    for (size_t p = 0; p < N; ++p)
      median->px[p] = (sdf[p] <= upper) && orImage->px[p];
    return median;
  }

  // X:556-793
  void Interpolate1to1(int axis, T label, IV i, IV j, const ConnPtr & iConn, T iId, const ConnPtr & jConn, T jId,
                       const IV translation[2], bool recursive)
  {
    ++nNodes;
    if (!recursive)
      ++nRoot;
    IV iTrans[2], jTrans[2];
    Rect iRegion = iConn->r, jRegion = jConn->r;
    IV jBottom[2];
    bool carry = false;
    for (int d = 0; d < 2; ++d)
    {
      if (!carry)
      {
        iTrans[d] = translation[d] / 2;
        carry = translation[d] % 2;
      }
      else if (translation[d] % 2 == 0)
        iTrans[d] = translation[d] / 2;
      else
      {
        iTrans[d] = translation[d] > 0 ? translation[d] / 2 + 1 : translation[d] / 2 - 1;
        carry = false;
      }
      jTrans[d] = iTrans[d] - translation[d];
      iRegion.idx[d] += iTrans[d];
      jRegion.idx[d] += jTrans[d];
      jBottom[d] = jRegion.idx[d] + jRegion.sz[d] - 1;
    }
    Rect newRegion = iRegion;
    expand_rect(newRegion, jRegion.idx);
    expand_rect(newRegion, jBottom);
    IV mid = (i + j + (carry ? 1 : 0)) / 2;

    ConnPtr iConnT = TranslateImage(*iConn, iTrans, newRegion);
    ConnPtr jConnT = TranslateImage(*jConn, jTrans, newRegion);
    if (!recursive)
    {
      newRegion = BoundingBox(*iConnT);
      Rect jBB = BoundingBox(*jConnT);
      IV i2[2] = { jBB.idx[0], jBB.idx[1] };
      expand_rect(newRegion, i2);
      for (int d = 0; d < 2; ++d)
        i2[d] += jBB.sz[d] - 1;
      expand_rect(newRegion, i2);
    }

    BoolPtr iSlice = std::make_shared<BoolSlice>(), jSlice = std::make_shared<BoolSlice>(),
            inter = std::make_shared<BoolSlice>();
    iSlice->alloc(newRegion);
    jSlice->alloc(newRegion);
    inter->alloc(newRegion);
    for (IV y = 0; y < newRegion.sz[1]; ++y)
      for (IV x = 0; x < newRegion.sz[0]; ++x)
      {
        IV gx = newRegion.idx[0] + x, gy = newRegion.idx[1] + y;
        bool a = iConnT->at(gx, gy) == iId, b = jConnT->at(gx, gy) == jId;
        size_t p = size_t(y) * size_t(newRegion.sz[0]) + size_t(x);
        iSlice->px[p] = a;
        jSlice->px[p] = b;
        inter->px[p] = a && b;
      }

    BoolPtr median =
      P.use_distance_transform ? FindMedianImageDistances(inter, iSlice, jSlice) : FindMedianImageDilations(inter, iSlice, jSlice);

    // clip to the output slice (X:679-712)
    int d0 = axis == 0 ? 1 : 0, d1 = axis == 2 ? 1 : 2;
    Rect sliceRegion;
    sliceRegion.idx[0] = 0;
    sliceRegion.idx[1] = 0;
    sliceRegion.sz[0] = dims[d0];
    sliceRegion.sz[1] = dims[d1];
    IV t0[2] = { 0, 0 };
    intersection_regions(t0, sliceRegion, newRegion);

    ConnPtr midConn = std::make_shared<ConnSlice>();
    midConn->alloc(newRegion);
    IdT written = 0;
    {
      std::lock_guard<std::mutex> hold(writeMutex);
      IV p[3];
      p[axis] = mid;
      for (IV y = 0; y < newRegion.sz[1]; ++y)
        for (IV x = 0; x < newRegion.sz[0]; ++x)
        {
          IV gx = newRegion.idx[0] + x, gy = newRegion.idx[1] + y;
          if (median->at(gx, gy))
          {
            midConn->at(gx, gy) = 1;
            p[d0] = gx;
            p[d1] = gy;
            T & o = out[(size_t(p[2]) * size_t(dims[1]) + size_t(p[1])) * size_t(dims[0]) + size_t(p[0])];
            if (o < label)
              o = label;
            ++written;
          }
        }
    }
    if (trace)
    {
      std::lock_guard<std::mutex> hold(traceMutex);
      fprintf(trace, "node axis=%d label=%ld i=%ld j=%ld mid=%ld iId=%ld jId=%ld t=(%ld,%ld) rec=%d region=(%ld,%ld,%ld,%ld) n=%lu\n",
              axis, long(label), i, j, mid, long(iId), long(jId), translation[0], translation[1], int(recursive), newRegion.idx[0],
              newRegion.idx[1], newRegion.sz[0], newRegion.sz[1], written);
    }

    if (std::labs(i - j) > 2)
    {
      bool first = std::labs(i - mid) > 1;
      bool second = std::labs(j - mid) > 1;
      if (first)
        Interpolate1to1(axis, label, i, mid, iConn, iId, midConn, 1, iTrans, true);
      if (second)
        Interpolate1to1(axis, label, j, mid, jConn, jId, midConn, 1, jTrans, true);
    }
  }

  // X:824-992
  void Interpolate1toN(int axis, T label, IV i, IV j, const ConnPtr & iConn, T iId, const ConnPtr & jConn, const PixelList & jIds,
                       const IV translation[2])
  {
    ++n1N;
    BoolPtr mask = std::make_shared<BoolSlice>();
    mask->alloc(iConn->r);
    for (size_t p = 0; p < mask->px.size(); ++p)
      mask->px[p] = (iConn->px[p] == iId);

    Rect iRegion = iConn->r, jRegion = jConn->r;
    std::vector<BoolPtr> blobs;
    for (size_t x = 0; x < jIds.size(); ++x)
    {
      BoolPtr b = std::make_shared<BoolSlice>();
      b->alloc(iConn->r);
      blobs.push_back(b);
    }
    ConnSlice belongs;
    belongs.alloc(iConn->r);
    intersection_regions(translation, iRegion, jRegion);
    for (IV y = 0; y < iRegion.sz[1]; ++y)
      for (IV x = 0; x < iRegion.sz[0]; ++x)
      {
        IV ix = iRegion.idx[0] + x, iy = iRegion.idx[1] + y;
        if (mask->at(ix, iy))
        {
          T jVal = jConn->at(jRegion.idx[0] + x, jRegion.idx[1] + y);
          auto res = std::find(jIds.begin(), jIds.end(), jVal);
          if (res != jIds.end())
          {
            blobs[res - jIds.begin()]->at(ix, iy) = 1;
            belongs.at(ix, iy) = T(res - jIds.begin() + 1);
          }
        }
      }

    bool hollowedMaskEmpty;
    const size_t N = mask->px.size();
    do
    {
      for (size_t x = 0; x < jIds.size(); ++x)
        blobs[x] = Dilate1(blobs[x], mask);
      hollowedMaskEmpty = true;
      for (size_t p = 0; p < N; ++p)
      {
        if (!mask->px[p])
          continue;
        if (!belongs.px[p])
        {
          size_t x = 0;
          while (x < jIds.size() && !blobs[x]->px[p])
            ++x;
          if (x < jIds.size())
          {
            belongs.px[p] = T(x + 1);
            for (x++; x < jIds.size(); x++)
              blobs[x]->px[p] = 0;
          }
          else
            hollowedMaskEmpty = false;
        }
        else
        {
          for (size_t x = 0; x < jIds.size(); ++x)
            if (size_t(belongs.px[p]) != x + 1)
              blobs[x]->px[p] = 0;
        }
      }
    } while (!hollowedMaskEmpty);
    blobs.clear();

    std::vector<ConnPtr> conns;
    for (size_t x = 0; x < jIds.size(); ++x)
    {
      ConnPtr c = std::make_shared<ConnSlice>();
      c->alloc(iConn->r);
      conns.push_back(c);
    }
    for (size_t p = 0; p < N; ++p)
      if (belongs.px[p] > 0)
        conns[size_t(belongs.px[p]) - 1]->px[p] = iId;
    for (size_t x = 0; x < jIds.size(); ++x)
      Interpolate1to1(axis, label, i, j, conns[x], iId, jConn, jIds[x], translation, false);
  }

  // X:203-251
  void Extrapolate(int axis, T label, IV i, IV j, const ConnPtr & iConn, T iId)
  {
    ++nExtra;
    PixelList ids(1, iId);
    IV c[2];
    Centroid(*iConn, ids, c);
    Rect reg3;
    reg3.sz[0] = reg3.sz[1] = 3;
    reg3.idx[0] = c[0] - 1;
    reg3.idx[1] = c[1] - 1;
    ConnPtr ph = std::make_shared<ConnSlice>();
    ph->alloc(reg3);
    ph->at(c[0], c[1]) = iId;
    IV t[2];
    Align(*iConn, iId, *ph, ids, t);
    // the region is re-indexed without moving the buffer: the phantom pixel travels with it
    ph->r.idx[0] -= t[0];
    ph->r.idx[1] -= t[1];
    IV t0[2] = { 0, 0 };
    if (trace)
    {
      std::lock_guard<std::mutex> hold(traceMutex);
      fprintf(trace, "extrap axis=%d label=%ld i=%ld j=%ld id=%ld c=(%ld,%ld) t=(%ld,%ld)\n", axis, long(label), i, j, long(iId), c[0], c[1],
              t[0], t[1]);
    }
    Interpolate1to1(axis, label, i, j, iConn, iId, ph, iId, t0, false);
  }

  void traceAlign(const char * what, int axis, T label, IV i, IV j, T id, const PixelList & ids, const IV t[2])
  {
    if (!trace)
      return;
    std::lock_guard<std::mutex> hold(traceMutex);
    fprintf(trace, "%s axis=%d label=%ld i=%ld j=%ld id=%ld ids=[", what, axis, long(label), i, j, long(id));
    for (T v : ids)
      fprintf(trace, "%ld,", long(v));
    fprintf(trace, "] t=(%ld,%ld)\n", t[0], t[1]);
  }

  // X:1240-1447
  void InterpolateBetweenTwo(int axis, T label, IV i, IV j, const ConnPtr & iconn, const ConnPtr & jconn)
  {
    typedef std::set<std::pair<T, T>> PairSet;
    PairSet pairs, unwantedPairs, uncleanPairs;
    for (size_t p = 0; p < iconn->px.size(); ++p)
    {
      T a = iconn->px[p], b = jconn->px[p];
      if (a != 0 || b != 0)
      {
        uncleanPairs.insert(std::make_pair(a, b));
        if (a != 0 && b != 0)
        {
          unwantedPairs.insert(std::make_pair(T(0), b));
          unwantedPairs.insert(std::make_pair(a, T(0)));
        }
      }
    }
    std::set_difference(uncleanPairs.begin(), uncleanPairs.end(), unwantedPairs.begin(), unwantedPairs.end(),
                        std::inserter(pairs, pairs.end()));

    auto p = pairs.begin();
    while (p != pairs.end())
    {
      if (p->second == 0)
      {
        if (P.use_extrapolation)
          Extrapolate(axis, label, i, j, iconn, p->first);
        pairs.erase(p++);
      }
      else if (p->first == 0)
      {
        if (P.use_extrapolation)
          Extrapolate(axis, label, j, i, jconn, p->second);
        pairs.erase(p++);
      }
      else
        ++p;
    }

    std::map<T, IdT> iCounts, jCounts;
    for (p = pairs.begin(); p != pairs.end(); ++p)
    {
      iCounts[p->first]++;
      jCounts[p->second]++;
    }

    p = pairs.begin();
    while (p != pairs.end())
    {
      if (iCounts[p->first] == 1 && jCounts[p->second] == 1)
      {
        PixelList ids(1, p->second);
        IV t[2];
        Align(*iconn, p->first, *jconn, ids, t);
        traceAlign("1to1", axis, label, i, j, p->first, ids, t);
        Interpolate1to1(axis, label, i, j, iconn, p->first, jconn, p->second, t, false);
        iCounts.erase(p->first);
        jCounts.erase(p->second);
        pairs.erase(p++);
      }
      else
        ++p;
    }

    PixelList ids;
    p = pairs.begin();
    while (p != pairs.end())
    {
      ids.clear();
      if (iCounts[p->first] == 1) // M-to-1
      {
        for (auto rest = pairs.begin(); rest != pairs.end(); ++rest)
          if (rest->second == p->second)
            ids.push_back(rest->first);
        IV t[2];
        Align(*jconn, p->second, *iconn, ids, t);
        traceAlign("Mto1", axis, label, j, i, p->second, ids, t);
        Interpolate1toN(axis, label, j, i, jconn, p->second, iconn, ids, t);
        auto rest = pairs.begin();
        while (rest != pairs.end())
        {
          if (rest != p && rest->second == p->second)
          {
            --iCounts[rest->first];
            --jCounts[rest->second];
            pairs.erase(rest++);
          }
          else
            ++rest;
        }
        --iCounts[p->first];
        --jCounts[p->second];
        pairs.erase(p++);
      }
      else if (jCounts[p->second] == 1) // 1-to-N
      {
        for (auto rest = pairs.begin(); rest != pairs.end(); ++rest)
          if (rest->first == p->first)
            ids.push_back(rest->second);
        IV t[2];
        Align(*iconn, p->first, *jconn, ids, t);
        traceAlign("1toN", axis, label, i, j, p->first, ids, t);
        Interpolate1toN(axis, label, i, j, iconn, p->first, jconn, ids, t);
        auto rest = pairs.begin();
        ++rest; // X:1387-1388: the scan starts at the second element
        while (rest != pairs.end())
        {
          if (rest != p && rest->first == p->first)
          {
            --iCounts[rest->first];
            --jCounts[rest->second];
            pairs.erase(rest++);
          }
          else
            ++rest;
        }
        --iCounts[p->first];
        --jCounts[p->second];
        pairs.erase(p++);
      }
      else
        ++p;
    }

    p = pairs.begin();
    while (p != pairs.end())
    {
      ids.clear();
      for (auto rest = p; rest != pairs.end(); ++rest)
        if (rest->first == p->first)
          ids.push_back(rest->second);
      IV t[2];
      Align(*iconn, p->first, *jconn, ids, t);
      traceAlign("MtoN", axis, label, i, j, p->first, ids, t);
      Interpolate1toN(axis, label, i, j, iconn, p->first, jconn, ids, t);
      auto rest = p;
      ++rest;
      while (rest != pairs.end())
      {
        if (rest->first == p->first)
          pairs.erase(rest++);
        else
          ++rest;
      }
      pairs.erase(p++);
    }
  }

  struct Segment
  {
    int axis;
    T label;
    IV i, j;
    ConnPtr iconn, jconn;
  };

  double tCC = 0, tInterp = 0;
  int64_t nSegments = 0;
  int64_t globalSegment = 0;
  bool keepSegment()
  {
    int64_t k = globalSegment++;
    return P.shard_count <= 1 || (k % P.shard_count) == P.shard_index;
  }

  // X:1449-1539
  void InterpolateAlong(int axis)
  {
    auto c0 = std::chrono::steady_clock::now();
    std::vector<Segment> segments;
    for (auto it = labeledSlices[axis].begin(); it != labeledSlices[axis].end(); ++it)
    {
      if (P.label == 0 || T(P.label) == it->first)
      {
        auto prev = it->second.begin();
        if (prev == it->second.end())
          continue;
        Box3 ri;
        for (int a = 0; a < 3; ++a)
        {
          ri.idx[a] = 0;
          ri.sz[a] = dims[a];
        }
        if (!P.use_custom_slice_positions)
        {
          auto bb = bboxes.find(it->first);
          if (bb == bboxes.end())
            continue;
          ri = bb->second;
        }
        ri.sz[axis] = 0;
        ri.idx[axis] = *prev;
        IdT cnt;
        ConnPtr iconn = RegionedConnectedComponents(ri, axis, it->first, cnt);
        auto next = it->second.begin();
        for (++next; next != it->second.end(); ++next)
        {
          Box3 rj = ri;
          rj.idx[axis] = *next;
          ConnPtr jconn = RegionedConnectedComponents(rj, axis, it->first, cnt);
          if (*prev + 1 < *next && keepSegment())
          {
            Segment s;
            s.axis = axis;
            s.label = it->first;
            s.i = *prev;
            s.j = *next;
            s.iconn = iconn;
            s.jconn = jconn;
            segments.push_back(s);
          }
          iconn = jconn;
          prev = next;
        }
      }
    }
    auto c1 = std::chrono::steady_clock::now();
    nSegments += int64_t(segments.size());

    // X:1523-1537 ParallelizeArray over segments
    int nt = std::max(1, P.threads);
    std::atomic<size_t> nextSeg{ 0 };
    std::atomic<bool> failed{ false };
    std::string err;
    std::mutex errMutex;
    auto worker = [&]() {
      for (;;)
      {
        size_t k = nextSeg.fetch_add(1);
        if (k >= segments.size() || failed.load())
          return;
        try
        {
          Segment & s = segments[k];
          InterpolateBetweenTwo(s.axis, s.label, s.i, s.j, s.iconn, s.jconn);
        }
        catch (std::exception & e)
        {
          std::lock_guard<std::mutex> hold(errMutex);
          failed = true;
          err = e.what();
        }
      }
    };
    if (nt == 1)
      worker();
    else
    {
      std::vector<std::thread> pool;
      for (int k = 0; k < nt; ++k)
        pool.emplace_back(worker);
      for (auto & th : pool)
        th.join();
    }
    auto c2 = std::chrono::steady_clock::now();
    tCC += std::chrono::duration<double>(c1 - c0).count();
    tInterp += std::chrono::duration<double>(c2 - c1).count();
    if (failed)
      throw std::runtime_error(err);
  }

  // X:1559-1637
  void GenerateData(Stats * st, const std::vector<std::map<T, std::set<IV>>> * custom)
  {
    auto c0 = std::chrono::steady_clock::now();
    const size_t N = size_t(dims[0]) * size_t(dims[1]) * size_t(dims[2]);
    std::memset(out, 0, N * sizeof(T)); // AllocateOutputs: zero-filled (X:1541-1557)
    DetermineSliceOrientations();
    if (P.use_custom_slice_positions && custom)
      labeledSlices = *custom;
    auto c1 = std::chrono::steady_clock::now();
    double tFinal = 0;
    if (bboxes.empty() && !P.use_custom_slice_positions)
    {
      std::memcpy(out, in, N * sizeof(T));
    }
    else
    {
      if (P.axis == -1)
      {
        bool aggregate[3] = { false, false, false };
        for (int a = 0; a < 3; ++a)
        {
          if (P.label == 0)
          {
            // X:1593-1599: the loop's net effect is "some label has more than one slice on this axis"
            for (auto & kv : labeledSlices[a])
              if (kv.second.size() > 1)
                aggregate[a] = true;
          }
          else
          {
            auto f = labeledSlices[a].find(T(P.label));
            if (f != labeledSlices[a].end() && f->second.size() > 1)
              aggregate[a] = true;
          }
        }
        for (int a = 0; a < 3; ++a)
          if (aggregate[a])
            InterpolateAlong(a);
      }
      else
        InterpolateAlong(P.axis);
      auto f0 = std::chrono::steady_clock::now();
      for (size_t p = 0; p < N; ++p)
        if (in[p] != 0)
          out[p] = in[p];
      tFinal = std::chrono::duration<double>(std::chrono::steady_clock::now() - f0).count();
    }
    auto c2 = std::chrono::steady_clock::now();
    if (st)
    {
      st->segments = nSegments;
      st->root_jobs = nRoot;
      st->nodes = nNodes;
      st->align_calls = nAlign;
      st->align_evals = nEval;
      st->extrapolations = nExtra;
      st->one_to_n = n1N;
      st->t_detect = std::chrono::duration<double>(c1 - c0).count();
      st->t_cc = tCC;
      st->t_interp = tInterp;
      st->t_final = tFinal;
      st->t_total = std::chrono::duration<double>(c2 - c0).count();
    }
  }
};

static thread_local std::string g_err;

template <typename T>
static int
run_typed(const void * in, void * out, const int64_t * dims, const Params * p, Stats * st, const int64_t * custom, int64_t ncustom)
{
  Oracle<T> o;
  o.in = static_cast<const T *>(in);
  o.out = static_cast<T *>(out);
  for (int a = 0; a < 3; ++a)
    o.dims[a] = IV(dims[a]);
  o.P = *p;
  const char * tr = std::getenv("MCI_ORACLE_TRACE");
  if (tr && *tr)
    o.trace = std::fopen(tr, "w");
  std::vector<std::map<T, std::set<IV>>> cust(3);
  for (int64_t k = 0; k < ncustom; ++k) // triplets (axis, label, slice)
    cust[size_t(custom[3 * k])][T(custom[3 * k + 1])].insert(IV(custom[3 * k + 2]));
  int rc = 0;
  try
  {
    o.GenerateData(st, p->use_custom_slice_positions ? &cust : nullptr);
  }
  catch (std::exception & e)
  {
    g_err = e.what();
    rc = 1;
  }
  if (o.trace)
    std::fclose(o.trace);
  return rc;
}

} // namespace

extern "C"
{
  // ptype: 0 = uint8, 1 = uint16, 2 = int16
  int mci_oracle_run(const void * in, void * out, const int64_t dims[3], int ptype, const Params * p, Stats * st,
                     const int64_t * custom_triplets, int64_t n_custom)
  {
    Stats local{};
    if (!st)
      st = &local;
    std::memset(st, 0, sizeof(Stats));
    switch (ptype)
    {
      case 0:
        return run_typed<uint8_t>(in, out, dims, p, st, custom_triplets, n_custom);
      case 1:
        return run_typed<uint16_t>(in, out, dims, p, st, custom_triplets, n_custom);
      case 2:
        return run_typed<int16_t>(in, out, dims, p, st, custom_triplets, n_custom);
    }
    g_err = "unsupported pixel type";
    return 2;
  }

  const char * mci_oracle_last_error() { return g_err.c_str(); }

  // slice detection only: writes triplets (axis,label,slice) and per-label bboxes (label, idx[3], size[3])
  int64_t mci_oracle_detect(const void * in, const int64_t dims[3], int ptype, int axis, int64_t * triplets, int64_t cap_triplets,
                            int64_t * boxes, int64_t cap_boxes, int64_t * n_boxes)
  {
    if (ptype != 1)
      return -1;
    Oracle<uint16_t> o;
    o.in = static_cast<const uint16_t *>(in);
    for (int a = 0; a < 3; ++a)
      o.dims[a] = IV(dims[a]);
    o.P.axis = axis;
    o.DetermineSliceOrientations();
    int64_t n = 0;
    for (int a = 0; a < 3; ++a)
      for (auto & kv : o.labeledSlices[a])
        for (IV s : kv.second)
        {
          if (n < cap_triplets)
          {
            triplets[3 * n] = a;
            triplets[3 * n + 1] = kv.first;
            triplets[3 * n + 2] = s;
          }
          ++n;
        }
    int64_t nb = 0;
    for (auto & kv : o.bboxes)
    {
      if (nb < cap_boxes)
      {
        boxes[7 * nb] = kv.first;
        for (int a = 0; a < 3; ++a)
        {
          boxes[7 * nb + 1 + a] = kv.second.idx[a];
          boxes[7 * nb + 4 + a] = kv.second.sz[a];
        }
      }
      ++nb;
    }
    *n_boxes = nb;
    return n;
  }

  // primitives, exposed so the tests can pin them against scipy / brute force
  int64_t mci_oracle_cc2d(const uint8_t * mask, int64_t w, int64_t h, int fully, uint16_t * ids)
  {
    BoolSlice m;
    Rect r = { { 0, 0 }, { IV(w), IV(h) } };
    m.alloc(r);
    std::memcpy(m.px.data(), mask, size_t(w * h));
    Slice<uint16_t> o;
    int64_t n = int64_t(connected_components<uint16_t>(m, fully != 0, o));
    std::memcpy(ids, o.px.data(), size_t(w * h) * 2);
    return n;
  }

  void mci_oracle_edt2(const uint8_t * mask, int64_t w, int64_t h, int brute, int64_t * d2)
  {
    BoolSlice m;
    Rect r = { { 0, 0 }, { IV(w), IV(h) } };
    m.alloc(r);
    std::memcpy(m.px.data(), mask, size_t(w * h));
    std::vector<int64_t> o;
    if (brute)
      edt_squared_brute(m, o);
    else
      edt_squared(m, o);
    std::memcpy(d2, o.data(), size_t(w * h) * 8);
  }

  void mci_oracle_dilate(const uint8_t * mask, int64_t w, int64_t h, int ball, uint8_t * o)
  {
    BoolSlice m;
    Rect r = { { 0, 0 }, { IV(w), IV(h) } };
    m.alloc(r);
    std::memcpy(m.px.data(), mask, size_t(w * h));
    BoolPtr d = dilate_r1(m, ball != 0);
    std::memcpy(o, d->px.data(), size_t(w * h));
  }

  // median of two shapes already in a common frame (no alignment): exposes FindMedianImage{Distances,Dilations}
  int mci_oracle_median(const uint8_t * iMask, const uint8_t * jMask, int64_t w, int64_t h, int ptype, int use_dt, int ball, uint8_t * o)
  {
    Rect r = { { 0, 0 }, { IV(w), IV(h) } };
    BoolPtr a = std::make_shared<BoolSlice>(), b = std::make_shared<BoolSlice>(), x = std::make_shared<BoolSlice>();
    a->alloc(r);
    b->alloc(r);
    x->alloc(r);
    for (int64_t p = 0; p < w * h; ++p)
    {
      a->px[p] = iMask[p] != 0;
      b->px[p] = jMask[p] != 0;
      x->px[p] = a->px[p] && b->px[p];
    }
    try
    {
      BoolPtr m;
      Params P{};
      P.use_ball = ball;
      if (ptype == 0)
      {
        Oracle<uint8_t> orc;
        orc.P = P;
        m = use_dt ? orc.FindMedianImageDistances(x, a, b) : orc.FindMedianImageDilations(x, a, b);
      }
      else if (ptype == 1)
      {
        Oracle<uint16_t> orc;
        orc.P = P;
        m = use_dt ? orc.FindMedianImageDistances(x, a, b) : orc.FindMedianImageDilations(x, a, b);
      }
      else
      {
        Oracle<int16_t> orc;
        orc.P = P;
        m = use_dt ? orc.FindMedianImageDistances(x, a, b) : orc.FindMedianImageDilations(x, a, b);
      }
      std::memcpy(o, m->px.data(), size_t(w * h));
    }
    catch (std::exception & e)
    {
      g_err = e.what();
      return 1;
    }
    return 0;
  }
}
