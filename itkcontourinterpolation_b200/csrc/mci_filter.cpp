// mci_filter.cpp -- host scheduler (layer 2 of include/mci_b200.h).
//
// Restates the HOST side of itk::MorphologicalContourInterpolator: GenerateData (X:1559-1637),
// InterpolateAlong's segment list (X:1449-1521) and InterpolateBetweenTwo's pair bookkeeping
// (X:1250-1446), where X = reference include/itkMorphologicalContourInterpolator.hxx.  It calls
// CUDA only through the phase-level C ABI; it never touches a voxel itself.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <thread>
#include <atomic>
#include <utility>
#include <vector>

#include "../../include/mci_b200.h"

// NVTX ranges per phase of a run: visible in Nsight Systems / ncu --nvtx; free when no tool is attached
#include <nvtx3/nvToolsExt.h>

namespace
{

struct NvtxRange
{
  explicit NvtxRange(const char * name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

struct Box
{
  int idx[3], size[3];
};

struct FilterState
{
  std::vector<std::map<int64_t, std::set<int>>> labeledSlices;
  std::map<int64_t, Box> bboxes;
};

std::mutex g_mutex;
std::map<mci_ctx *, FilterState> g_states;

FilterState &
state_of(mci_ctx * ctx)
{
  std::lock_guard<std::mutex> hold(g_mutex);
  return g_states[ctx];
}

double
now_ms()
{
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Segment
{
  int axis;
  int64_t label;
  int i, j;   // slice positions
  int si, sj; // indices into the slice batch
};

typedef std::pair<int, int> IdPair;

// InterpolateBetweenTwo's set algebra (X:1274-1446) for one segment; emits jobs instead of calling.
void
schedule_segment(const Segment & seg, const std::set<IdPair> & uncleanPairs, bool useExtrapolation, std::vector<mci_job> & jobs,
                 std::vector<int32_t> & bIds)
{
  std::set<IdPair> unwantedPairs, pairs;
  for (const IdPair & q : uncleanPairs)
    if (q.first != 0 && q.second != 0)
    {
      unwantedPairs.insert(IdPair(0, q.second));
      unwantedPairs.insert(IdPair(q.first, 0));
    }
  for (const IdPair & q : uncleanPairs)
    if (!unwantedPairs.count(q))
      pairs.insert(q);

  auto emit = [&](int kind, bool swapped, int aId, const std::vector<int> & ids) {
    mci_job jb;
    jb.kind = kind;
    jb.slice_a = swapped ? seg.sj : seg.si;
    jb.a_id = aId;
    jb.slice_b = swapped ? seg.si : seg.sj;
    jb.n_b = (int)ids.size();
    jb.b_ids_offset = (int)bIds.size();
    jb.pos_a = swapped ? seg.j : seg.i;
    jb.pos_b = swapped ? seg.i : seg.j;
    for (int v : ids)
      bIds.push_back(v);
    jobs.push_back(jb);
  };
  const std::vector<int> none;

  // components without overlap: extrapolate (X:1280-1304)
  auto p = pairs.begin();
  while (p != pairs.end())
  {
    if (p->second == 0)
    {
      if (useExtrapolation)
        emit(MCI_JOB_EXTRAPOLATE, false, p->first, none);
      pairs.erase(p++);
    }
    else if (p->first == 0)
    {
      if (useExtrapolation)
        emit(MCI_JOB_EXTRAPOLATE, true, p->second, none);
      pairs.erase(p++);
    }
    else
      ++p;
  }

  std::map<int, long> iCounts, jCounts;
  for (p = pairs.begin(); p != pairs.end(); ++p)
  {
    iCounts[p->first]++;
    jCounts[p->second]++;
  }

  // 1-to-1 (X:1315-1333)
  p = pairs.begin();
  while (p != pairs.end())
  {
    if (iCounts[p->first] == 1 && jCounts[p->second] == 1)
    {
      emit(MCI_JOB_ONE_TO_ONE, false, p->first, std::vector<int>(1, p->second));
      iCounts.erase(p->first);
      jCounts.erase(p->second);
      pairs.erase(p++);
    }
    else
      ++p;
  }

  // M-to-1 and 1-to-N (X:1335-1411)
  std::vector<int> ids;
  p = pairs.begin();
  while (p != pairs.end())
  {
    ids.clear();
    if (iCounts[p->first] == 1) // M-to-1: roles swap
    {
      for (auto rest = pairs.begin(); rest != pairs.end(); ++rest)
        if (rest->second == p->second)
          ids.push_back(rest->first);
      emit(MCI_JOB_ONE_TO_N, true, p->second, ids);
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
      emit(MCI_JOB_ONE_TO_N, false, p->first, ids);
      auto rest = pairs.begin();
      ++rest; // the reference's erase scan starts at the second element (X:1387-1388)
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

  // M-to-N: one 1-to-N per remaining i component (X:1413-1446)
  p = pairs.begin();
  while (p != pairs.end())
  {
    ids.clear();
    for (auto rest = p; rest != pairs.end(); ++rest)
      if (rest->first == p->first)
        ids.push_back(rest->second);
    emit(MCI_JOB_ONE_TO_N, false, p->first, ids);
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

// The same bookkeeping for the segments that need none of the set algebra -- after the unwanted pairs are dropped,
// every component takes part in at most one pair (extrapolations and 1-to-1 correspondences only; all but a few
// segments of a real volume) -- on the sorted pair list itself, without a std::set / std::map node allocation per
// pair.  q[0..n) sorted, distinct.  Emits exactly what schedule_segment emits (extrapolations in set order
// X:1280-1304, then the 1-to-1 jobs in set order X:1315-1333) or returns false without emitting anything.
bool
schedule_simple(const Segment & seg, const IdPair * q, int n, bool useExtrapolation, std::vector<mci_job> & jobs,
                std::vector<int32_t> & bIds)
{
  if (n > 32)
    return false;
  unsigned keep = 0; // bit k: q[k] survives the set difference of X:1274-1278
  for (int k = 0; k < n; ++k)
  {
    bool unwanted = false;
    if (q[k].first == 0 || q[k].second == 0)
      for (int m = 0; m < n && !unwanted; ++m)
        if (q[m].first != 0 && q[m].second != 0 && (q[k].first == 0 ? q[m].second == q[k].second : q[m].first == q[k].first))
          unwanted = true;
    if (!unwanted)
      keep |= 1u << k;
  }
  for (int k = 0; k < n; ++k)
    if (((keep >> k) & 1u) && q[k].first != 0 && q[k].second != 0)
      for (int m = k + 1; m < n; ++m)
        if (q[m].first != 0 && q[m].second != 0 && (q[m].first == q[k].first || q[m].second == q[k].second))
          return false; // 1-to-N, M-to-1 or M-to-N: the literal path
  auto emit = [&](int kind, bool swapped, int aId, int nB, int bId) {
    mci_job jb;
    jb.kind = kind;
    jb.slice_a = swapped ? seg.sj : seg.si;
    jb.a_id = aId;
    jb.slice_b = swapped ? seg.si : seg.sj;
    jb.n_b = nB;
    jb.b_ids_offset = (int)bIds.size();
    jb.pos_a = swapped ? seg.j : seg.i;
    jb.pos_b = swapped ? seg.i : seg.j;
    if (nB)
      bIds.push_back(bId);
    jobs.push_back(jb);
  };
  if (useExtrapolation)
    for (int k = 0; k < n; ++k)
      if ((keep >> k) & 1u)
      {
        if (q[k].second == 0)
          emit(MCI_JOB_EXTRAPOLATE, false, q[k].first, 0, 0);
        else if (q[k].first == 0)
          emit(MCI_JOB_EXTRAPOLATE, true, q[k].second, 0, 0);
      }
  for (int k = 0; k < n; ++k)
    if (((keep >> k) & 1u) && q[k].first != 0 && q[k].second != 0)
      emit(MCI_JOB_ONE_TO_ONE, false, q[k].first, 1, q[k].second);
  return true;
}

// one segment's sorted, distinct pairs -> jobs
void
schedule_pairs_of_segment(const Segment & seg, const IdPair * q, int n, bool useExtrapolation, std::vector<mci_job> & jobs,
                          std::vector<int32_t> & bIds)
{
  if (schedule_simple(seg, q, n, useExtrapolation, jobs, bIds))
    return;
  schedule_segment(seg, std::set<IdPair>(q, q + n), useExtrapolation, jobs, bIds);
}

int
detect_into_state(mci_ctx * ctx, const mci_filter_options * opt, FilterState & st)
{
  int32_t nl = 0, ns = 0;
  int rc = mci_detect_slices(ctx, opt->axis, &nl, &ns);
  if (rc)
    return rc;
  std::vector<mci_label_info> labels(nl);
  std::vector<mci_labeled_slice> slices(ns);
  rc = mci_get_detection(ctx, labels.data(), slices.data());
  if (rc)
    return rc;
  st.labeledSlices.assign(3, std::map<int64_t, std::set<int>>());
  st.bboxes.clear();
  for (const mci_label_info & L : labels)
  {
    Box b;
    for (int a = 0; a < 3; ++a)
    {
      b.idx[a] = L.bbox_index[a];
      b.size[a] = L.bbox_size[a];
    }
    st.bboxes[L.label] = b;
  }
  for (const mci_labeled_slice & s : slices)
    st.labeledSlices[s.axis][s.label].insert(s.slice);
  return MCI_OK;
}

int
run_from_state(mci_ctx * ctx, const mci_filter_options * opt, const int64_t dims[3], FilterState & st, int keepLo, int keepHi,
               mci_filter_stats * stats, double t1);

// GenerateData (X:1559-1637) on a volume already resident on the device.
int
run_core(mci_ctx * ctx, const mci_filter_options * opt, const int64_t dims[3], const int64_t * custom, int64_t nCustom,
         mci_filter_stats * stats)
{
  FilterState & st = state_of(ctx);
  double t0 = now_ms();
  int rc;
  {
    NvtxRange r("mci.detect_slices");
    rc = detect_into_state(ctx, opt, st); // also initialises the output with a copy of the input
  }
  if (rc)
    return rc;
  if (opt->use_custom_slice_positions)
  {
    // bounding boxes come from the detection pass, positions from the caller (X:1567-1572)
    st.labeledSlices.assign(3, std::map<int64_t, std::set<int>>());
    for (int64_t k = 0; k < nCustom; ++k)
    {
      int64_t a = custom[3 * k];
      if (a < 0 || a > 2)
        return MCI_ERR_ARG;
      st.labeledSlices[(size_t)a][custom[3 * k + 1]].insert((int)custom[3 * k + 2]);
    }
  }
  double t1 = now_ms();
  stats->ms_detect = t1 - t0;
  return run_from_state(ctx, opt, dims, st, 0, -1, stats, t1);
}

// Everything of GenerateData after slice detection.  keepLo <= keepHi restricts the run to the segments with a mid
// slice inside [keepLo, keepHi) along their axis (slab-sharded runs: the context holds a z-range of the volume).
int
run_from_state(mci_ctx * ctx, const mci_filter_options * opt, const int64_t dims[3], FilterState & st, int keepLo, int keepHi,
               mci_filter_stats * stats, double t1)
{
  int rc = MCI_OK;
  stats->n_labels = (int64_t)st.bboxes.size();
  for (int a = 0; a < 3; ++a)
    for (auto & kv : st.labeledSlices[a])
      stats->n_annotated_slices += (int64_t)kv.second.size();
  if (st.bboxes.empty() && !opt->use_custom_slice_positions)
    return MCI_OK; // no contours: output == input (X:1578-1583)

  // which axes (X:1585-1621)
  bool doAxis[3] = { false, false, false };
  if (opt->axis == -1)
  {
    for (int a = 0; a < 3; ++a)
    {
      if (opt->label == 0)
      {
        for (auto & kv : st.labeledSlices[a])
          if (kv.second.size() > 1)
            doAxis[a] = true;
      }
      else
      {
        auto f = st.labeledSlices[a].find(opt->label);
        if (f != st.labeledSlices[a].end() && f->second.size() > 1)
          doAxis[a] = true;
      }
    }
  }
  else
    doAxis[opt->axis] = true;

  // segment list of InterpolateAlong (X:1456-1521), all axes batched: every axis reads the INPUT volume
  // and results combine with the max rule, so axes are independent
  std::vector<mci_slice_desc> sliceDescs;
  std::vector<Segment> segments;
  for (int axis = 0; axis < 3; ++axis)
  {
    if (!doAxis[axis])
      continue;
    const int d0 = axis == 0 ? 1 : 0, d1 = axis == 2 ? 1 : 2;
    for (auto & kv : st.labeledSlices[axis])
    {
      if (!(opt->label == 0 || opt->label == kv.first) || kv.second.empty())
        continue;
      Box region;
      for (int a = 0; a < 3; ++a)
      {
        region.idx[a] = 0;
        region.size[a] = (int)dims[a];
      }
      if (!opt->use_custom_slice_positions)
      {
        auto bb = st.bboxes.find(kv.first);
        if (bb == st.bboxes.end())
          continue;
        region = bb->second;
      }
      // position -> batch index, created on demand.  Positions are visited in ascending order and a slice is shared
      // only by two consecutive segments (j of one, i of the next), so remembering the last one created is enough
      int lastPos = -1, lastIdx = -1;
      auto sliceOf = [&](int pos) {
        if (pos == lastPos)
          return lastIdx;
        mci_slice_desc d;
        d.axis = axis;
        d.slice = pos;
        d.label = kv.first;
        d.u0 = region.idx[d0];
        d.v0 = region.idx[d1];
        d.w = region.size[d0];
        d.h = region.size[d1];
        sliceDescs.push_back(d);
        lastPos = pos;
        lastIdx = (int)sliceDescs.size() - 1;
        return lastIdx;
      };
      std::vector<int> positions(kv.second.begin(), kv.second.end());
      for (int pos : positions)
        if (pos < 0 || pos >= dims[axis])
          return MCI_ERR_ARG;
      for (size_t k = 1; k < positions.size(); ++k)
      {
        if (positions[k - 1] + 1 < positions[k]) // only if they are not adjacent slices (X:1502)
        {
          if (keepLo <= keepHi && (positions[k - 1] + 1 >= keepHi || positions[k] - 1 < keepLo))
            continue; // slab-sharded run: no mid slice of this segment lies in the owned range
          Segment s;
          s.axis = axis;
          s.label = kv.first;
          s.i = positions[k - 1];
          s.j = positions[k];
          s.si = sliceOf(s.i);
          s.sj = sliceOf(s.j);
          segments.push_back(s);
        }
      }
    }
  }
  if (opt->shard_count > 1)
  {
    // multi-GPU: keep every shard_count-th segment; slices no kept segment touches are dropped from the batch
    if (opt->shard_index < 0 || opt->shard_index >= opt->shard_count)
      return MCI_ERR_ARG;
    std::vector<Segment> mine;
    std::vector<mci_slice_desc> keptDescs;
    std::map<int, int> remap;
    for (size_t k = 0; k < segments.size(); ++k)
    {
      if ((int)(k % (size_t)opt->shard_count) != opt->shard_index)
        continue;
      Segment s = segments[k];
      for (int * idx : { &s.si, &s.sj })
      {
        auto f = remap.find(*idx);
        if (f == remap.end())
        {
          keptDescs.push_back(sliceDescs[(size_t)*idx]);
          f = remap.insert(std::make_pair(*idx, (int)keptDescs.size() - 1)).first;
        }
        *idx = f->second;
      }
      mine.push_back(s);
    }
    segments.swap(mine);
    sliceDescs.swap(keptDescs);
  }
  stats->n_segments = (int64_t)segments.size();
  if (segments.empty())
    return MCI_OK;

  // K1/K2/K3: component images of every annotated slice that bounds a segment, then the pair scan of every
  // segment, launched back to back: one host round trip brings back counts, pairs and component statistics
  std::vector<int32_t> ncomp(sliceDescs.size());
  std::vector<mci_segment_desc> segDescs(segments.size());
  for (size_t k = 0; k < segments.size(); ++k)
  {
    segDescs[k].slice_i = segments[k].si;
    segDescs[k].slice_j = segments[k].sj;
  }
  int32_t nPairs = 0;
  nvtxRangePushA("mci.components_and_pairs");
  rc = mci_label_and_match(ctx, (int32_t)sliceDescs.size(), sliceDescs.data(), opt->use_ball_structuring_element,
                           (int32_t)segDescs.size(), segDescs.data(), ncomp.data(), &nPairs);
  nvtxRangePop();
  if (rc)
    return rc;
  for (int32_t c : ncomp)
    stats->n_components += c;
  double t2 = now_ms();
  stats->ms_components = t2 - t1;
  std::vector<mci_pair> pairs(nPairs);
  rc = mci_get_pairs(ctx, pairs.data());
  if (rc)
    return rc;
  stats->n_pairs = nPairs;
  // the pairs of each segment as a sorted, duplicate-free range (std::set iteration order) of one array
  std::vector<int> segStart(segments.size() + 1, 0);
  for (const mci_pair & q : pairs)
    ++segStart[(size_t)q.segment + 1];
  for (size_t k = 0; k < segments.size(); ++k)
    segStart[k + 1] += segStart[k];
  std::vector<IdPair> sorted(pairs.size());
  {
    std::vector<int> fill(segStart.begin(), segStart.end() - 1);
    for (const mci_pair & q : pairs)
      sorted[(size_t)fill[(size_t)q.segment]++] = IdPair(q.i_id, q.j_id);
  }
  double t3 = now_ms();
  stats->ms_pairs = t3 - t2;

  // pair bookkeeping -> jobs
  std::vector<mci_job> jobs;
  std::vector<int32_t> bIds;
  jobs.reserve(pairs.size());
  bIds.reserve(pairs.size());
  {
    NvtxRange r("mci.schedule_pairs");
    for (size_t k = 0; k < segments.size(); ++k)
    {
      IdPair * b = sorted.data() + segStart[k], * e = sorted.data() + segStart[k + 1];
      std::sort(b, e);
      e = std::unique(b, e);
      schedule_pairs_of_segment(segments[k], b, (int)(e - b), opt->use_extrapolation != 0, jobs, bIds);
    }
  }
  stats->n_jobs = (int64_t)jobs.size();

  mci_params prm;
  prm.heuristic_alignment = opt->heuristic_alignment;
  prm.use_distance_transform = opt->use_distance_transform;
  prm.use_ball = opt->use_ball_structuring_element;
  prm.min_align_iters = 8;   // pow(2, ImageDimension) with ImageDimension = 3 (X:87)
  prm.max_align_iters = 216; // pow(6, ImageDimension)                         (X:89)
  {
    NvtxRange r("mci.interpolate");
    rc = mci_interpolate(ctx, (int32_t)jobs.size(), jobs.data(), bIds.data(), (int32_t)bIds.size(), &prm);
  }
  if (rc)
    return rc;
  return MCI_OK;
}

} // namespace

extern "C"
{

  void mci_filter_default_options(mci_filter_options * opt)
  {
    if (!opt)
      return;
    opt->label = 0;
    opt->axis = -1;
    opt->heuristic_alignment = 1;
    opt->use_distance_transform = 1;
    opt->use_ball_structuring_element = 0;
    opt->use_custom_slice_positions = 0;
    opt->use_extrapolation = 1;
    opt->shard_index = 0;
    opt->shard_count = 1;
  }

  int mci_filter_run(mci_ctx * ctx, const mci_filter_options * opt, const void * host_in, void * host_out, const int64_t dims[3],
                     int pixel_type, const int64_t * custom, int64_t nCustom, mci_filter_stats * stats)
  {
    if (!ctx || !opt || !host_in || !host_out || !dims || opt->axis < -1 || opt->axis > 2)
      return MCI_ERR_ARG;
    mci_filter_stats local;
    if (!stats)
      stats = &local;
    std::memset(stats, 0, sizeof(*stats));
    int64_t l0 = mci_launch_count(ctx);
    double t0 = now_ms();
    int rc = mci_upload_volume(ctx, host_in, dims, pixel_type);
    if (rc)
      return rc;
    double t1 = now_ms();
    stats->ms_upload = t1 - t0;
    rc = run_core(ctx, opt, dims, custom, nCustom, stats);
    if (rc)
      return rc;
    double t2 = now_ms();
    rc = mci_download_volume(ctx, host_out); // synchronises and surfaces device-side errors
    double t3 = now_ms();
    stats->ms_interpolate = t3 - t1 - stats->ms_detect - stats->ms_components - stats->ms_pairs;
    stats->ms_download = t3 - t2;
    stats->ms_total = t3 - t0;
    stats->n_launches = mci_launch_count(ctx) - l0;
    return rc;
  }

  int mci_filter_run_resident(mci_ctx * ctx, const mci_filter_options * opt, const int64_t * custom, int64_t nCustom,
                              mci_filter_stats * stats)
  {
    if (!ctx || !opt || opt->axis < -1 || opt->axis > 2)
      return MCI_ERR_ARG;
    mci_filter_stats local;
    if (!stats)
      stats = &local;
    std::memset(stats, 0, sizeof(*stats));
    int64_t l0 = mci_launch_count(ctx);
    double t0 = now_ms();
    int64_t dims[3];
    int ptype = 0;
    int rc = mci_volume_dims(ctx, dims, &ptype);
    if (rc)
      return rc;
    rc = run_core(ctx, opt, dims, custom, nCustom, stats);
    if (rc)
      return rc;
    rc = mci_synchronize(ctx);
    double t1 = now_ms();
    stats->ms_interpolate = t1 - t0 - stats->ms_detect - stats->ms_components - stats->ms_pairs;
    stats->ms_total = t1 - t0;
    stats->n_launches = mci_launch_count(ctx) - l0;
    return rc;
  }

  // Slab-sharded runs: the context holds a contiguous range of the volume along opt->axis (its own slab plus the
  // annotated slices that bound the segments reaching into it); slice sets and bounding boxes are the merged tables
  // of all ranks, already shifted into the range's coordinates.  See include/mci_b200.h.
  int mci_filter_run_with_detection(mci_ctx * ctx, const mci_filter_options * opt, const mci_label_info * labels, int32_t nLabels,
                                    const mci_labeled_slice * slices, int32_t nSlices, int32_t keepLo, int32_t keepHi,
                                    mci_filter_stats * stats)
  {
    if (!ctx || !opt || opt->axis < -1 || opt->axis > 2 || nLabels < 0 || nSlices < 0 || (nLabels && !labels) || (nSlices && !slices))
      return MCI_ERR_ARG;
    mci_filter_stats local;
    if (!stats)
      stats = &local;
    std::memset(stats, 0, sizeof(*stats));
    int64_t l0 = mci_launch_count(ctx);
    double t0 = now_ms();
    int64_t dims[3];
    int ptype = 0;
    int rc = mci_volume_dims(ctx, dims, &ptype);
    if (rc)
      return rc;
    FilterState & st = state_of(ctx);
    st.labeledSlices.assign(3, std::map<int64_t, std::set<int>>());
    st.bboxes.clear();
    for (int32_t k = 0; k < nLabels; ++k)
    {
      Box b;
      for (int a = 0; a < 3; ++a)
      {
        b.idx[a] = labels[k].bbox_index[a];
        b.size[a] = labels[k].bbox_size[a];
        if (b.idx[a] < 0 || b.size[a] < 1)
          return MCI_ERR_ARG;
      }
      st.bboxes[labels[k].label] = b;
    }
    for (int32_t k = 0; k < nSlices; ++k)
    {
      if (slices[k].axis < 0 || slices[k].axis > 2)
        return MCI_ERR_ARG;
      if (opt->axis == -1 || opt->axis == slices[k].axis)
        st.labeledSlices[(size_t)slices[k].axis][slices[k].label].insert(slices[k].slice);
    }
    rc = run_from_state(ctx, opt, dims, st, keepLo, keepHi, stats, t0);
    if (rc)
      return rc;
    rc = mci_synchronize(ctx);
    double t1 = now_ms();
    stats->ms_interpolate = t1 - t0 - stats->ms_components - stats->ms_pairs;
    stats->ms_total = t1 - t0;
    stats->n_launches = mci_launch_count(ctx) - l0;
    return rc;
  }

  int mci_filter_run_batch(mci_ctx * const * ctxs, int32_t nCtx, const mci_filter_options * opt, const void * const * hostIn,
                           void * const * hostOut, int32_t nVolumes, const int64_t dims[3], int pixelType, mci_filter_stats * stats)
  {
    if (!ctxs || nCtx < 1 || !opt || nVolumes < 0 || (nVolumes > 0 && (!hostIn || !hostOut)) || !dims)
      return MCI_ERR_ARG;
    for (int32_t c = 0; c < nCtx; ++c)
      if (!ctxs[c])
        return MCI_ERR_ARG;
    // one worker per context; volumes are handed out in order, so with n_ctx in flight the copy engines of both
    // directions and the SMs are busy at the same time.  (What makes this work: the small descriptor / result tables
    // of a run never enter a copy-engine queue -- mci_cabi.cu, TableCopy -- so a context in its compute phase is not
    // held up behind another context's 100+ MB volume copy.)
    std::atomic<int32_t> next(0);
    std::atomic<int> firstError(MCI_OK);
    auto worker = [&](int32_t c) {
      for (;;)
      {
        int32_t k = next.fetch_add(1);
        if (k >= nVolumes || firstError.load() != MCI_OK)
          return;
        int rc = mci_filter_run(ctxs[c], opt, hostIn[k], hostOut[k], dims, pixelType, nullptr, 0, stats ? stats + k : nullptr);
        if (rc)
        {
          int expected = MCI_OK;
          firstError.compare_exchange_strong(expected, rc);
          return;
        }
      }
    };
    const int32_t nWorkers = nCtx < nVolumes ? nCtx : nVolumes;
    for (int32_t c = 0; c < nWorkers; ++c)
      mci_set_wait_mode(ctxs[c], nWorkers > 1); // several host threads: sleep through the bulk copies instead of spinning
    std::vector<std::thread> threads;
    for (int32_t c = 1; c < nWorkers; ++c)
      threads.emplace_back(worker, c);
    if (nWorkers > 0)
      worker(0);
    for (std::thread & t : threads)
      t.join();
    for (int32_t c = 0; c < nWorkers; ++c)
      mci_set_wait_mode(ctxs[c], 0);
    return firstError.load();
  }

  int mci_filter_detect(mci_ctx * ctx, const mci_filter_options * opt, int64_t * nTriplets)
  {
    if (!ctx || !opt)
      return MCI_ERR_ARG;
    FilterState & st = state_of(ctx);
    int rc = detect_into_state(ctx, opt, st);
    if (rc)
      return rc;
    int64_t n = 0;
    for (int a = 0; a < 3; ++a)
      for (auto & kv : st.labeledSlices[a])
        n += (int64_t)kv.second.size();
    if (nTriplets)
      *nTriplets = n;
    return MCI_OK;
  }

  int mci_filter_get_labeled_slices(mci_ctx * ctx, int64_t * triplets)
  {
    if (!ctx || !triplets)
      return MCI_ERR_ARG;
    FilterState & st = state_of(ctx);
    int64_t n = 0;
    for (int a = 0; a < (int)st.labeledSlices.size(); ++a)
      for (auto & kv : st.labeledSlices[a])
        for (int s : kv.second)
        {
          triplets[3 * n] = a;
          triplets[3 * n + 1] = kv.first;
          triplets[3 * n + 2] = s;
          ++n;
        }
    return MCI_OK;
  }

  // The pair bookkeeping of InterpolateBetweenTwo (X:1274-1446) for ONE segment, exposed so that it can be checked
  // without a GPU: `pairs` are the distinct (iId, jId) != (0,0) of the segment as mci_match_pairs emits them.  Jobs
  // come back with slice_a/slice_b = 0 for the i slice and 1 for the j slice, pos_a/pos_b = 0 / 1 likewise.
  int mci_filter_schedule_pairs(const mci_pair * pairs, int32_t n_pairs, int use_extrapolation, mci_job * jobs, int32_t cap_jobs,
                                int32_t * n_jobs, int32_t * b_ids, int32_t cap_b_ids, int32_t * n_b_ids)
  {
    if ((n_pairs > 0 && !pairs) || !n_jobs || !n_b_ids)
      return MCI_ERR_ARG;
    Segment seg;
    seg.axis = 0;
    seg.label = 1;
    seg.i = 0;
    seg.j = 1;
    seg.si = 0;
    seg.sj = 1;
    std::vector<IdPair> unclean;
    for (int32_t k = 0; k < n_pairs; ++k)
      unclean.push_back(IdPair(pairs[k].i_id, pairs[k].j_id));
    std::sort(unclean.begin(), unclean.end());
    unclean.erase(std::unique(unclean.begin(), unclean.end()), unclean.end());
    std::vector<mci_job> js;
    std::vector<int32_t> ids;
    if (use_extrapolation & 2) // bit 1 (tests): the literal set algebra for every pair set, simple or not
      schedule_segment(seg, std::set<IdPair>(unclean.begin(), unclean.end()), (use_extrapolation & 1) != 0, js, ids);
    else
      schedule_pairs_of_segment(seg, unclean.data(), (int)unclean.size(), (use_extrapolation & 1) != 0, js, ids);
    *n_jobs = (int32_t)js.size();
    *n_b_ids = (int32_t)ids.size();
    if ((int32_t)js.size() > cap_jobs || (int32_t)ids.size() > cap_b_ids)
      return MCI_ERR_CAPACITY;
    for (size_t k = 0; k < js.size(); ++k)
      jobs[k] = js[k];
    for (size_t k = 0; k < ids.size(); ++k)
      b_ids[k] = ids[k];
    return MCI_OK;
  }

  void mci_filter_forget(mci_ctx * ctx)
  {
    std::lock_guard<std::mutex> hold(g_mutex);
    g_states.erase(ctx);
  }
}
