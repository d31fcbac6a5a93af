// mci_launch.h -- host-callable launch wrappers of the kernels in mci_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "mci_types.h"
#include "../../include/mci_b200.h"

namespace mci
{

struct Launch
{
  cudaStream_t stream;
  int sms;      // SM count of the device (148 on B200)
  int maxSmem;  // opt-in dynamic shared memory per CTA
  int64_t count; // kernels launched so far
  const char * failedKernel; // first launch that returned an error, if any
  cudaError_t failedCode;
  // optional per-launch timing with CUDA events on the launching stream (bench.py roofline)
  bool profile;
  struct Rec
  {
    const char * name;
    cudaEvent_t a, b;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t pendingStart;
  cudaEvent_t take()
  {
    cudaEvent_t e;
    if (!pool.empty())
    {
      e = pool.back();
      pool.pop_back();
    }
    else
      cudaEventCreate(&e);
    return e;
  }
  // opt-in dynamic shared memory is a per-device function attribute: every context sets it once for itself (contexts
  // may live on different devices and be driven by different host threads, so no process-wide flag)
  unsigned attrDone = 0;
  template <typename F>
  void ensure_max_smem(F fn, int bit, int bytes)
  {
    if (!((attrDone >> bit) & 1u))
    {
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      attrDone |= 1u << bit;
    }
  }
  const char * only = nullptr; // when set, only launches of this kernel are timed
  bool pending = false;
  static bool same(const char * a, const char * b)
  {
    while (*a && *a == *b)
    {
      ++a;
      ++b;
    }
    return *a == *b;
  }
  void pre(const char * name)
  {
    pending = profile && (!only || same(only, name));
    if (pending)
    {
      pendingStart = take();
      cudaEventRecord(pendingStart, stream);
    }
  }
  void post(const char * name)
  {
    ++count;
    if (pending)
    {
      pending = false;
      Rec r;
      r.name = name;
      r.a = pendingStart;
      r.b = take();
      cudaEventRecord(r.b, stream);
      recs.push_back(r);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess && !failedKernel)
    {
      failedKernel = name;
      failedCode = e;
    }
  }
};

// Programmatic dependent launch: the kernel is queued with programmatic stream serialization, so its launch is
// processed while its predecessor in the stream still runs; it calls pdl_wait() before touching anything the
// predecessor wrote (no early trigger is used: its CTAs start when the predecessor has completed and flushed).
// Shaves the launch gap off chains of short dependent kernels (the six kernels of a distance-transform level).
template <typename... KArgs, typename... Args>
inline void
launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args)
{
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
__device__ __forceinline__ void
pdl_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

struct DtLevel
{
  const Node * nodes;
  const int * count;
  int maxNodes;
  NodeWork * work;
  uint32_t * arena;
  unsigned long long * arenaTop;
  unsigned long long arenaCap;
  uint32_t * scratch;
  unsigned long long * scratchTop;
  unsigned long long scratchCap;
  DtItem * items;
  int * itemCount;
  DtItem * witems;
  int * witemCount;
  int itemCap;
  DtItem * hitems; // histogram work items (8-word chunks holding xor pixels)
  int * hitemCount;
  int hitemCap;
  unsigned * hist;
  int histStride, histSide;
  unsigned * bigHist;
  int * bigCount;
  int nbig;
  Node * next;
  int * nextCount;
  int nextCap;
  EmitRec * emit; // list of produced mid slices (for the multi-GPU combine)
  int * emitCount;
  int emitCap;
  unsigned long long emitBase; // arena word where mid shapes start
  const void * in;
  void * out;
  long long dims[3];
  int * err;
};
void launch_dt_level(Launch & L, int ptype, const DtLevel & lv);

// the whole recursion of a root node in one CTA (k_dt_tree, mci_dt.cu)
struct TreeParams
{
  Node * roots;    // node queue: roots first, then room for every child
  const int * rootCount;
  int * rootNext;  // [0] head, [1] wrapped-bin tables in use, [2] tail (children), [3] pushed, [4] done; zeroed before the launch
  int * flags;     // one per child slot: 1 = published (zeroed before the launch)
  int queueCap;    // child slots
  uint32_t * arena;
  unsigned long long * arenaTop;
  unsigned long long arenaCap;
  uint32_t * scratch;
  unsigned long long scratchPerCta; // words
  int histSide;
  unsigned * bigHist;
  int * bigCount;
  int nbig;
  EmitRec * emit;
  int * emitCount;
  int emitCap;
  unsigned long long emitBase;
  const void * in;
  void * out;
  long long VX, VY, VZ;
  int composeAxis2; // 1: slices perpendicular to z and y are written by the compose passes that follow, not by the nodes
  int * err;
};
void launch_dt_tree(Launch & L, int ptype, const TreeParams & P, int nRootsMax, int grid, int threads);
// volume writes of all mid slices perpendicular to `axis` (1 or 2), no atomics (tables: 7 * dims[axis] + 1 + cap ints,
// prepared by the caller: counts / starts / fills zero, box minima INT_MAX-like, maxima INT_MIN-like)
bool compose_tables_in_one_launch(int cap); // few mid slices: one table kernel that also initialises the tables
void launch_compose(Launch & L, int ptype, int axis, const EmitRec * recs, const int * nrecsDev, int cap, const uint32_t * words,
                    int * tables, const void * in, void * out, const long long dims[3]);

long long median_tile_count(const long long dims[3]);
void launch_median3d(Launch & L, int ptype, const void * in, void * out, const long long dims[3], int radius, unsigned * summary);
void launch_combine_max(Launch & L, int ptype, void * dst, const void * src, long long n);
void launch_detect(Launch & L, int ptype, const void * in, void * out, const long long dims[3], int axisSel, int * bbox, int nl,
                   uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2, uint32_t * chunkBits);
void launch_compact_labels(Launch & L, const int * bbox, int nl, const uint32_t * sliceBits, int wordsPerLabel, int wo1, int wo2,
                           int signedLabels, mci_label_info * labels, mci_labeled_slice * slices, int capSlices, int * counters);
void launch_cc(Launch & L, int ptype, const void * in, const SliceDev * slices, int nslices, const RowTile * tiles, int ntiles,
               uint32_t * bits, uint32_t * wordRun, uint32_t * runComp, uint32_t * runScratch, uint16_t * conn, int * ncomp,
               int * compBase, int * total, CompStat * stats, int statCap, int fully, int maxRows, int * err);
void launch_pairs(Launch & L, const SliceDev * slices, const SegDev * segs, const RowTile * tiles, int ntiles, const uint16_t * conn,
                  u64 * table, uint32_t tableMask, mci_pair * pairs, int capPairs, int * nPairs, int * err);
void launch_extract(Launch & L, const SliceDev * slices, const ExtractJob * jobs, int njobs, const uint16_t * conn, uint32_t * arena);
void launch_align(Launch & L, const AlignJob * jobs, int njobs, const MaskRef * refs, const uint32_t * arena, uint32_t * gscratch,
                  int heuristic, int2 * result, int4 * dbg, int smemBytes, int * err);
void launch_split(Launch & L, const SplitJob * jobs, int njobs, const MaskRef * refs, uint32_t * arena, uint32_t * gscratch,
                  const int2 * trans, int ball, int maxWords, int * err);
void launch_make_roots(Launch & L, const RootJob * roots, int nroots, MaskRef * refs, uint32_t * arena, const int2 * trans,
                       Node * nodes);
void launch_nodes(Launch & L, int ptype, const Node * nodes, const int * count, int maxNodes, Node * next, int * nextCount,
                  int nextCap, uint32_t * arena, unsigned long long * arenaTop, unsigned long long arenaCap, const void * in,
                  void * out, const long long dims[3], int useDT, int ball, unsigned * bigHist, uint32_t * slab,
                  unsigned long long slabWords, unsigned short * arr, unsigned long long arrPixels, int grid, int smemBytes,
                  EmitRec * emit, int * emitCount, int emitCap, unsigned long long emitBase, int * err);
void launch_apply_masks(Launch & L, int ptype, const EmitRec * recs, int nrecs, const int * nrecsDev, const uint32_t * words,
                        const void * in, void * out, const long long dims[3], int skipAxis0);
void launch_apply_axis0(Launch & L, int ptype, const Axis0Group * groups, int ngroups, const EmitRec * recs, const int * slots,
                        const uint32_t * words, const void * in, void * out, const long long dims[3]);

} // namespace mci
