// mci_cabi.cu -- context and phase-level C ABI (layer 1 of include/mci_b200.h).
// Host code here only marshals descriptors, sizes device arenas and launches kernels; every
// pixel of the hot path is touched on the device.  There is no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "mci_launch.h"

using namespace mci;

namespace
{

struct DBuf
{
  void * p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HBuf
{
  void * p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes)
  {
    if (bytes <= cap)
      return cudaSuccess;
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    // mapped: kernels read / write these buffers directly (see table_copy below)
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// host blob -> one H2D copy -> typed device pointers
struct Blob
{
  std::vector<uint8_t> bytes;
  template <typename T>
  size_t add(const std::vector<T> & v)
  {
    size_t off = (bytes.size() + 255) & ~size_t(255);
    bytes.resize(off + std::max<size_t>(v.size() * sizeof(T), 1));
    if (!v.empty())
      std::memcpy(bytes.data() + off, v.data(), v.size() * sizeof(T));
    return off;
  }
  size_t reserve(size_t n)
  {
    size_t off = (bytes.size() + 255) & ~size_t(255);
    bytes.resize(off + std::max<size_t>(n, 1));
    return off;
  }
};

} // namespace

struct mci_ctx
{
  int device = 0;
  Launch L{};
  std::string err;
  // volume
  int ptype = -1;
  long long dims[3] = { 0, 0, 0 };
  size_t nvox = 0;
  void * dIn = nullptr;
  void * dOut = nullptr;
  DBuf ownIn, ownOut, ownTmp, dMedTiles;
  // detection
  DBuf dBBox, dSliceBits, dLabels, dLabeledSlices, dCounters, dChunkBits;
  HBuf hErr;
  HBuf hStage, hStage2, hRes; // hRes: pinned landing zone of the small device->host result tables
  int resLabels = 0, resSlices = 0, resPairs = 0, resStats = 0; // how many entries of each table hRes holds
  int nLabels = 0, nLabeledSlices = 0;
  bool detected = false;
  // slices / components
  std::vector<mci_slice_desc> slices;
  std::vector<SliceDev> sliceDev;
  std::vector<int> ncomp, compBase;
  int totalComp = 0;
  bool statsOnHost = false;
  std::vector<CompStat> hStats;
  DBuf dSliceBlob, dBits, dWordRun, dRunComp, dRunScratch, dConn, dNcomp, dCompBase, dTotal, dStats;
  SliceDev * dSlices = nullptr;
  int statCap = 0;
  bool deferSync = false; // set by mci_label_and_match: the two phases launch back to back, one round trip at the end
  // pairs
  DBuf dSegBlob, dPairTable, dPairs;
  int nPairs = 0, capPairs = 0;
  // interpolation
  DBuf dJobBlob, dArena, dScratch, dBigHist, dSlab, dNodes, dLevelCounts, dArenaTop, dTrans;
  DBuf dWork, dItems, dLevelScratch, dHist, dDtBig, dDtCounters, dEmit, dArrival, dCompose;
  unsigned long long emitBase = 0; // arena word where the mid shapes of the last run start
  int emitCap = 0;
  bool haveEmit = false;
  int axisJobs[3] = { 0, 0, 0 }; // jobs of the current mci_interpolate call per slice axis
  bool histClean = false, dtBigClean = false;
  int bigHistGrid = 0;
  int nJobs = 0;
  // second stream: 1-to-N splits run beside the alignment of the other jobs
  cudaStream_t stream2 = nullptr;
  cudaEvent_t evAlignedSplits = nullptr, evSplitDone = nullptr;
  // bulk volume copies are waited for with a BLOCKING event (the host thread sleeps for the milliseconds they take
  // instead of spinning: several contexts / ranks share the host cores); everything else spin-waits for latency
  cudaEvent_t evBulk = nullptr;
  bool bulkPending = false;
  bool sleepOnBulk = false; // mci_set_wait_mode
  // error flag
  int * dErr = nullptr;
};

// layout of the pinned result cache (mci_ctx::hRes)
static const int kSpecLabels = 1024, kSpecSlices = 16384, kSpecPairs = 16384, kSpecStats = 8192, kSpecNcomp = 65536;
static const size_t kResLabelsOff = 64;
static const size_t kResSlicesOff = kResLabelsOff + size_t(kSpecLabels) * sizeof(mci_label_info);
static const size_t kResNcompOff = kResSlicesOff + size_t(kSpecSlices) * sizeof(mci_labeled_slice);
static const size_t kResPairsOff = kResNcompOff + size_t(2 * kSpecNcomp) * sizeof(int);
static const size_t kResStatsOff = kResPairsOff + size_t(kSpecPairs) * sizeof(mci_pair);
static const size_t kResBytes = kResStatsOff + size_t(kSpecStats) * sizeof(mci::CompStat);

#define CK(call)                                                                                                               \
  do                                                                                                                           \
  {                                                                                                                            \
    cudaError_t e_ = (call);                                                                                                   \
    if (e_ != cudaSuccess)                                                                                                     \
    {                                                                                                                          \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                                           \
      return MCI_ERR_CUDA;                                                                                                     \
    }                                                                                                                          \
  } while (0)

// ---- small tables between host and device ------------------------------------------------------------------
// Descriptor blobs (host -> device) and result tables (device -> host) cross PCIe through MAPPED pinned memory with
// SM loads / stores, never through the copy engines: a copy engine serves its queue in submission order, so with
// several volumes in flight (mci_filter_run_batch) a 100-byte table copy would wait behind another context's
// 100+ MB volume copy, milliseconds each time.  Up to six segments per launch; a segment may be bounded by an element
// count that lives on the device, so speculative fixed-size transfers are not needed.
struct TableSeg
{
  uint32_t * dst;
  const uint32_t * src;
  unsigned nWords;          // fixed size, or the cap when count != nullptr
  const int * count;        // device pointer to an element count (or nullptr)
  unsigned wordsPerElem;
};
struct TableSegs
{
  TableSeg s[6];
  int n;
};
__global__ void __launch_bounds__(256) k_table_copy(TableSegs t)
{
  pdl_wait();
  unsigned n[6], nmax = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k)
  {
    n[k] = 0;
    if (k < t.n)
    {
      n[k] = t.s[k].nWords;
      if (t.s[k].count)
      {
        const long long want = (long long)max(*t.s[k].count, 0) * t.s[k].wordsPerElem;
        n[k] = (unsigned)min((long long)n[k], want);
      }
      nmax = max(nmax, n[k]);
    }
  }
  // the loads of every segment are issued before the first store: a load from mapped host memory is a PCIe round
  // trip, and the grid is sized for one word per thread and segment, so a launch costs one round trip, not one per
  // segment and grid stride
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nmax; i += gridDim.x * blockDim.x)
  {
    uint32_t v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (i < n[k])
        v[k] = t.s[k].src[i];
#pragma unroll
    for (int k = 0; k < 6; ++k)
      if (i < n[k])
        t.s[k].dst[i] = v[k];
  }
}
struct TableCopy
{
  TableSegs t;
  unsigned maxWords = 0;
  TableCopy() { t.n = 0; }
  // bytes must be a multiple of 4 (every table here is)
  void add(void * dst, const void * src, size_t bytes, const int * count = nullptr, unsigned bytesPerElem = 4)
  {
    TableSeg & g = t.s[t.n++];
    g.dst = (uint32_t *)dst;
    g.src = (const uint32_t *)src;
    g.nWords = (unsigned)((bytes + 3) / 4);
    g.count = count;
    g.wordsPerElem = bytesPerElem / 4;
    maxWords = std::max(maxWords, g.nWords);
  }
  void launch(Launch & L)
  {
    if (!t.n)
      return;
    const int grid = (int)std::max(1u, std::min((maxWords + 255) / 256, 1184u));
    L.pre("k_table_copy");
    launch_pdl(k_table_copy, grid, 256, 0, L.stream, t);
    L.post("k_table_copy");
  }
};

static int
fail(mci_ctx * ctx, int code, const std::string & msg)
{
  ctx->err = msg;
  return code;
}

// tc: result tables of the phase that ride along in the same launch as the error flag
static int
check_device_error(mci_ctx * ctx, TableCopy * riders = nullptr)
{
  CK(ctx->hErr.ensure(64));
  {
    TableCopy own;
    TableCopy & tc = riders ? *riders : own;
    tc.add(ctx->hErr.p, ctx->dErr, sizeof(int));
    tc.launch(ctx->L);
  }
  CK(cudaStreamSynchronize(ctx->L.stream));
  const int h = *(volatile int *)ctx->hErr.p;
  if (ctx->L.failedKernel)
  {
    ctx->err = std::string("launch of ") + ctx->L.failedKernel + " failed: " + cudaGetErrorString(ctx->L.failedCode);
    ctx->L.failedKernel = nullptr;
    return MCI_ERR_CUDA;
  }
  CK(cudaGetLastError());
  if (h == ERR_NONE)
    return MCI_OK;
  static const char * names[] = { "ok",
                                  "mask arena overflow",
                                  "node list overflow",
                                  "pair table overflow",
                                  "component table overflow",
                                  "component count or distance bin does not fit the pixel type",
                                  "alignment queue overflow",
                                  "node scratch too small",
                                  "1-to-N split cannot reach every mask pixel" };
  int z = 0;
  cudaMemcpyAsync(ctx->dErr, &z, sizeof(int), cudaMemcpyHostToDevice, ctx->L.stream);
  cudaStreamSynchronize(ctx->L.stream);
  ctx->histClean = false;
  ctx->dtBigClean = false;
  ctx->bigHistGrid = 0;
  ctx->err = std::string("device: ") + (h >= 0 && h <= 8 ? names[h] : "unknown");
  return h == ERR_PIXELTYPE ? MCI_ERR_PIXELTYPE : MCI_ERR_CAPACITY;
}

static void recycle_events(mci_ctx * ctx);

extern "C"
{

  int mci_create(mci_ctx ** out, int device)
  {
    if (!out)
      return MCI_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n)
      return MCI_ERR_NO_DEVICE; // no CPU fallback: the path needs a CUDA device
    mci_ctx * ctx = new mci_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->L.stream, cudaStreamNonBlocking) != cudaSuccess)
    {
      delete ctx;
      return MCI_ERR_CUDA;
    }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major != 10)
    {
      // the library holds sm_100a SASS only: on any other GPU every launch would fail with "no kernel image"
      cudaStreamDestroy(ctx->L.stream);
      delete ctx;
      return MCI_ERR_NO_DEVICE;
    }
    ctx->L.sms = prop.multiProcessorCount;
    ctx->L.maxSmem = (int)prop.sharedMemPerBlockOptin;
    ctx->L.count = 0;
    ctx->L.failedKernel = nullptr;
    ctx->L.failedCode = cudaSuccess;
    ctx->L.profile = false;
    if (cudaMalloc((void **)&ctx->dErr, sizeof(int)) != cudaSuccess)
    {
      delete ctx;
      return MCI_ERR_CUDA;
    }
    cudaMemset(ctx->dErr, 0, sizeof(int));
    cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&ctx->evAlignedSplits, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evSplitDone, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ctx->evBulk, cudaEventDisableTiming | cudaEventBlockingSync);
    *out = ctx;
    return MCI_OK;
  }

  void mci_destroy(mci_ctx * ctx)
  {
    if (!ctx)
      return;
    mci_filter_forget(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->L.stream);
    DBuf * bufs[] = { &ctx->ownTmp, &ctx->dMedTiles, &ctx->ownIn,     &ctx->ownOut,   &ctx->dBBox,     &ctx->dSliceBits, &ctx->dLabels,  &ctx->dLabeledSlices,
                      &ctx->dCounters, &ctx->dSliceBlob, &ctx->dBits, &ctx->dWordRun,    &ctx->dConn,    &ctx->dRunComp, &ctx->dRunScratch,
                      &ctx->dNcomp,    &ctx->dCompBase, &ctx->dTotal,   &ctx->dStats,     &ctx->dSegBlob, &ctx->dPairTable,
                      &ctx->dPairs,    &ctx->dJobBlob,  &ctx->dArena,   &ctx->dScratch,   &ctx->dBigHist, &ctx->dSlab,
                      &ctx->dNodes,    &ctx->dLevelCounts, &ctx->dArenaTop, &ctx->dTrans,
                      &ctx->dWork, &ctx->dItems, &ctx->dLevelScratch, &ctx->dHist, &ctx->dDtBig, &ctx->dDtCounters,
                      &ctx->dChunkBits, &ctx->dEmit, &ctx->dArrival, &ctx->dCompose };
    for (DBuf * b : bufs)
      b->release();
    recycle_events(ctx);
    for (cudaEvent_t e : ctx->L.pool)
      cudaEventDestroy(e);
    ctx->hStage.release();
    ctx->hErr.release();
    ctx->hRes.release();
    ctx->hStage2.release();
    cudaFree(ctx->dErr);
    if (ctx->stream2)
      cudaStreamDestroy(ctx->stream2);
    if (ctx->evAlignedSplits)
      cudaEventDestroy(ctx->evAlignedSplits);
    if (ctx->evSplitDone)
      cudaEventDestroy(ctx->evSplitDone);
    if (ctx->evBulk)
      cudaEventDestroy(ctx->evBulk);
    cudaStreamDestroy(ctx->L.stream);
    delete ctx;
  }

  const char * mci_last_error(const mci_ctx * ctx) { return ctx ? ctx->err.c_str() : "null context"; }

  void * mci_host_alloc(size_t bytes)
  {
    void * p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess)
      return nullptr;
    return p;
  }
  void mci_host_free(void * p)
  {
    if (p)
      cudaFreeHost(p);
  }
  int64_t mci_launch_count(const mci_ctx * ctx) { return ctx ? ctx->L.count : 0; }
  void * mci_stream(const mci_ctx * ctx) { return ctx ? (void *)ctx->L.stream : nullptr; }
  int mci_set_wait_mode(mci_ctx * ctx, int sleep_on_bulk_copies)
  {
    if (!ctx)
      return MCI_ERR_ARG;
    ctx->sleepOnBulk = sleep_on_bulk_copies != 0;
    ctx->bulkPending = false;
    return MCI_OK;
  }
  int mci_synchronize(mci_ctx * ctx)
  {
    if (!ctx)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    return check_device_error(ctx);
  }

} // extern C (helper below has C++ linkage)
static void recycle_events(mci_ctx * ctx)
  {
    for (auto & r : ctx->L.recs)
    {
      ctx->L.pool.push_back(r.a);
      ctx->L.pool.push_back(r.b);
    }
    ctx->L.recs.clear();
  }
extern "C"
{

  int mci_profile_enable(mci_ctx * ctx, int on)
  {
    if (!ctx)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->L.stream));
    recycle_events(ctx);
    ctx->L.profile = on != 0;
    ctx->L.only = on == 2 ? "k_copy_flag" : nullptr; // 2: time only the HBM-streaming kernel (negligible overhead)
    return MCI_OK;
  }

  int mci_profile_read(mci_ctx * ctx, mci_kernel_time * out, int32_t cap, int32_t * n)
  {
    if (!ctx || !n || (cap > 0 && !out))
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->L.stream));
    std::map<std::string, std::pair<int64_t, double>> agg;
    for (auto & r : ctx->L.recs)
    {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess)
      {
        auto & e = agg[r.name];
        e.first += 1;
        e.second += ms;
      }
    }
    recycle_events(ctx);
    int k = 0;
    for (auto & kv : agg)
    {
      if (k < cap)
      {
        std::memset(out[k].name, 0, sizeof(out[k].name));
        std::strncpy(out[k].name, kv.first.c_str(), sizeof(out[k].name) - 1);
        out[k].launches = kv.second.first;
        out[k].total_ms = kv.second.second;
      }
      ++k;
    }
    *n = k;
    return MCI_OK;
  }

  // ---- volume ---------------------------------------------------------------------------------
  static int set_dims(mci_ctx * ctx, const int64_t dims[3], int ptype)
  {
    if (!dims || ptype < MCI_U8 || ptype > MCI_I16)
      return fail(ctx, MCI_ERR_ARG, "bad dims / pixel type");
    for (int a = 0; a < 3; ++a)
      if (dims[a] < 1 || dims[a] > 8192)
        return fail(ctx, MCI_ERR_ARG, "each volume dimension must be in [1, 8192]");
    ctx->ptype = ptype;
    for (int a = 0; a < 3; ++a)
      ctx->dims[a] = dims[a];
    ctx->nvox = size_t(dims[0]) * size_t(dims[1]) * size_t(dims[2]);
    ctx->detected = false;
    return MCI_OK;
  }

  int mci_upload_volume(mci_ctx * ctx, const void * host, const int64_t dims[3], int ptype)
  {
    if (!ctx || !host)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = set_dims(ctx, dims, ptype);
    if (rc)
      return rc;
    size_t bytes = ctx->nvox * (ptype == MCI_U8 ? 1 : 2);
    CK(ctx->ownIn.ensure(bytes + 16));
    CK(ctx->ownOut.ensure(bytes + 16));
    ctx->dIn = ctx->ownIn.p;
    ctx->dOut = ctx->ownOut.p;
    CK(cudaMemcpyAsync(ctx->dIn, host, bytes, cudaMemcpyHostToDevice, ctx->L.stream));
    if (ctx->sleepOnBulk && bytes >= (size_t(8) << 20))
    {
      CK(cudaEventRecord(ctx->evBulk, ctx->L.stream));
      ctx->bulkPending = true;
    }
    return MCI_OK;
  }

  int mci_attach_device_volume(mci_ctx * ctx, const void * dev_in, void * dev_out, const int64_t dims[3], int ptype)
  {
    if (!ctx || !dev_in || !dev_out)
      return MCI_ERR_ARG;
    if ((reinterpret_cast<size_t>(dev_in) & 15) || (reinterpret_cast<size_t>(dev_out) & 15))
      return fail(ctx, MCI_ERR_ARG, "device volumes must be 16-byte aligned");
    int rc = set_dims(ctx, dims, ptype);
    if (rc)
      return rc;
    ctx->dIn = const_cast<void *>(dev_in);
    ctx->dOut = dev_out;
    return MCI_OK;
  }

  int mci_copy_input_range(mci_ctx * ctx, int64_t first, int64_t n)
  {
    if (!ctx || !ctx->dIn || !ctx->dOut || first < 0 || n < 0 || size_t(first) + size_t(n) > ctx->nvox)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const size_t s = ctx->ptype == MCI_U8 ? 1 : 2;
    if (n)
      CK(cudaMemcpyAsync((char *)ctx->dOut + size_t(first) * s, (const char *)ctx->dIn + size_t(first) * s, size_t(n) * s,
                         cudaMemcpyDeviceToDevice, ctx->L.stream));
    return MCI_OK;
  }

  int mci_download_volume(mci_ctx * ctx, void * host)
  {
    if (!ctx || !host || !ctx->dOut)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    size_t bytes = ctx->nvox * (ctx->ptype == MCI_U8 ? 1 : 2);
    CK(cudaMemcpyAsync(host, ctx->dOut, bytes, cudaMemcpyDeviceToHost, ctx->L.stream));
    if (ctx->sleepOnBulk && bytes >= (size_t(8) << 20))
    {
      CK(cudaEventRecord(ctx->evBulk, ctx->L.stream));
      CK(cudaEventSynchronize(ctx->evBulk)); // sleeps
    }
    ctx->bulkPending = false;
    return check_device_error(ctx);
  }

  void * mci_device_output(mci_ctx * ctx) { return ctx ? ctx->dOut : nullptr; }

  int mci_volume_dims(const mci_ctx * ctx, int64_t dims[3], int * pixel_type)
  {
    if (!ctx || !dims || ctx->ptype < 0)
      return MCI_ERR_STATE;
    for (int a = 0; a < 3; ++a)
      dims[a] = ctx->dims[a];
    if (pixel_type)
      *pixel_type = ctx->ptype;
    return MCI_OK;
  }

  int mci_combine_max(mci_ctx * ctx, void * dst, const void * src, int64_t n, int ptype)
  {
    if (!ctx || !dst || !src || n < 0 || ptype < MCI_U8 || ptype > MCI_I16)
      return MCI_ERR_ARG;
    if ((reinterpret_cast<size_t>(dst) & 15) || (reinterpret_cast<size_t>(src) & 15))
      return fail(ctx, MCI_ERR_ARG, "buffers must be 16-byte aligned");
    CK(cudaSetDevice(ctx->device));
    launch_combine_max(ctx->L, ptype, dst, src, n);
    return MCI_OK;
  }

  // ---- label smoothing after the interpolation (examples/mciExample.cxx:29-39) ----------------
  static int median_into(mci_ctx * ctx, const void * src, void * dst, int radius)
  {
    if (radius == 0)
    {
      CK(cudaMemcpyAsync(dst, src, ctx->nvox * (ctx->ptype == MCI_U8 ? 1 : 2), cudaMemcpyDeviceToDevice, ctx->L.stream));
      return MCI_OK;
    }
    CK(ctx->dMedTiles.ensure(size_t(median_tile_count(ctx->dims)) * sizeof(unsigned)));
    launch_median3d(ctx->L, ctx->ptype, src, dst, ctx->dims, radius, (unsigned *)ctx->dMedTiles.p);
    return MCI_OK;
  }

  int mci_median_filter_resident(mci_ctx * ctx, int radius)
  {
    if (!ctx || !ctx->dOut)
      return MCI_ERR_ARG;
    if (radius < 0 || radius > 3)
      return fail(ctx, MCI_ERR_ARG, "median filter radius must be in [0, 3]");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = ctx->nvox * (ctx->ptype == MCI_U8 ? 1 : 2);
    CK(ctx->ownTmp.ensure(bytes + 16));
    int rc = median_into(ctx, ctx->dOut, ctx->ownTmp.p, radius);
    if (rc)
      return rc;
    if (ctx->dOut == ctx->ownOut.p)
    {
      std::swap(ctx->ownOut, ctx->ownTmp); // the smoothed volume becomes the output; no copy
      ctx->dOut = ctx->ownOut.p;
    }
    else
      CK(cudaMemcpyAsync(ctx->dOut, ctx->ownTmp.p, bytes, cudaMemcpyDeviceToDevice, ctx->L.stream)); // caller-owned output
    ctx->haveEmit = false;
    return MCI_OK;
  }

  int mci_median_filter_run(mci_ctx * ctx, const void * host_in, void * host_out, const int64_t dims[3], int ptype, int radius)
  {
    if (!ctx || !host_in || !host_out)
      return MCI_ERR_ARG;
    if (radius < 0 || radius > 3)
      return fail(ctx, MCI_ERR_ARG, "median filter radius must be in [0, 3]");
    int rc = mci_upload_volume(ctx, host_in, dims, ptype);
    if (rc)
      return rc;
    rc = median_into(ctx, ctx->dIn, ctx->dOut, radius);
    if (rc)
      return rc;
    return mci_download_volume(ctx, host_out);
  }

  // ---- slice detection ------------------------------------------------------------------------
  int mci_detect_slices(mci_ctx * ctx, int axis, int32_t * n_labels, int32_t * n_labeled_slices)
  {
    if (!ctx || !ctx->dIn || axis < -1 || axis > 2)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const int nl = ctx->ptype == MCI_U8 ? 256 : 65536;
    int wpa[3];
    for (int a = 0; a < 3; ++a)
      wpa[a] = (int)((ctx->dims[a] + 31) / 32);
    const int wo1 = wpa[0], wo2 = wpa[0] + wpa[1], wpl = wpa[0] + wpa[1] + wpa[2];
    const int capSlices = 1 << 20;
    CK(ctx->dBBox.ensure(size_t(6) * nl * sizeof(int)));
    CK(ctx->dSliceBits.ensure(size_t(nl) * wpl * sizeof(uint32_t)));
    CK(ctx->dLabels.ensure(size_t(nl) * sizeof(mci_label_info)));
    CK(ctx->dLabeledSlices.ensure(size_t(capSlices) * sizeof(mci_labeled_slice)));
    CK(ctx->dCounters.ensure(2 * sizeof(int)));
    CK(ctx->dChunkBits.ensure((ctx->nvox / (ctx->ptype == MCI_U8 ? 16 : 8) / 32 + 2) * 4));
    cudaStream_t s = ctx->L.stream;
    CK(cudaMemsetAsync(ctx->dBBox.p, 0x7F, size_t(3) * nl * sizeof(int), s));
    CK(cudaMemsetAsync((int *)ctx->dBBox.p + size_t(3) * nl, 0xFF, size_t(3) * nl * sizeof(int), s));
    CK(cudaMemsetAsync(ctx->dSliceBits.p, 0, size_t(nl) * wpl * sizeof(uint32_t), s));
    CK(cudaMemsetAsync(ctx->dCounters.p, 0, 2 * sizeof(int), s));
    launch_detect(ctx->L, ctx->ptype, ctx->dIn, ctx->dOut, ctx->dims, axis, (int *)ctx->dBBox.p, nl, (uint32_t *)ctx->dSliceBits.p,
                  wpl, wo1, wo2, (uint32_t *)ctx->dChunkBits.p);
    launch_compact_labels(ctx->L, (int *)ctx->dBBox.p, nl, (uint32_t *)ctx->dSliceBits.p, wpl, wo1, wo2, ctx->ptype == MCI_I16,
                          (mci_label_info *)ctx->dLabels.p, (mci_labeled_slice *)ctx->dLabeledSlices.p, capSlices,
                          (int *)ctx->dCounters.p);
    // one round trip: the counters plus the first kSpecLabels / kSpecSlices table entries land in pinned memory
    // together (label volumes rarely hold more); mci_get_detection fetches the rest if there is more
    CK(ctx->hRes.ensure(kResBytes));
    uint8_t * hr = (uint8_t *)ctx->hRes.p;
    int * h = (int *)hr;
    {
      TableCopy tc;
      tc.add(h, ctx->dCounters.p, 2 * sizeof(int));
      tc.add(hr + kResLabelsOff, ctx->dLabels.p, size_t(std::min(nl, kSpecLabels)) * sizeof(mci_label_info), (int *)ctx->dCounters.p,
             sizeof(mci_label_info));
      tc.add(hr + kResSlicesOff, ctx->dLabeledSlices.p, size_t(kSpecSlices) * sizeof(mci_labeled_slice), (int *)ctx->dCounters.p + 1,
             sizeof(mci_labeled_slice));
      tc.launch(ctx->L);
    }
    if (ctx->bulkPending)
    {
      ctx->bulkPending = false;
      CK(cudaEventSynchronize(ctx->evBulk)); // the upload queued ahead of the detection kernels: sleep through it
    }
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (h[1] > capSlices)
      return fail(ctx, MCI_ERR_CAPACITY, "too many labelled slices");
    ctx->nLabels = h[0];
    ctx->nLabeledSlices = h[1];
    ctx->resLabels = std::min(h[0], std::min(nl, kSpecLabels));
    ctx->resSlices = std::min(h[1], kSpecSlices);
    ctx->detected = true;
    if (n_labels)
      *n_labels = h[0];
    if (n_labeled_slices)
      *n_labeled_slices = h[1];
    return MCI_OK;
  }

  int mci_get_detection(mci_ctx * ctx, mci_label_info * labels, mci_labeled_slice * slices)
  {
    if (!ctx || !ctx->detected)
      return MCI_ERR_STATE;
    CK(cudaSetDevice(ctx->device));
    const uint8_t * hr = (const uint8_t *)ctx->hRes.p;
    bool pending = false;
    if (labels && ctx->nLabels)
    {
      if (ctx->resLabels >= ctx->nLabels)
        std::memcpy(labels, hr + kResLabelsOff, size_t(ctx->nLabels) * sizeof(mci_label_info));
      else
      {
        CK(cudaMemcpyAsync(labels, ctx->dLabels.p, size_t(ctx->nLabels) * sizeof(mci_label_info), cudaMemcpyDeviceToHost,
                           ctx->L.stream));
        pending = true;
      }
    }
    if (slices && ctx->nLabeledSlices)
    {
      if (ctx->resSlices >= ctx->nLabeledSlices)
        std::memcpy(slices, hr + kResSlicesOff, size_t(ctx->nLabeledSlices) * sizeof(mci_labeled_slice));
      else
      {
        CK(cudaMemcpyAsync(slices, ctx->dLabeledSlices.p, size_t(ctx->nLabeledSlices) * sizeof(mci_labeled_slice),
                           cudaMemcpyDeviceToHost, ctx->L.stream));
        pending = true;
      }
    }
    if (pending)
      CK(cudaStreamSynchronize(ctx->L.stream));
    return MCI_OK;
  }

  // ---- connected components ---------------------------------------------------------------------
  static void make_tiles(const std::vector<int> & widths, const std::vector<int> & heights, std::vector<RowTile> & tiles)
  {
    tiles.clear();
    for (size_t k = 0; k < widths.size(); ++k)
    {
      int rows = std::max(1, 8192 / std::max(1, widths[k]));
      for (int r0 = 0; r0 < heights[k]; r0 += rows)
      {
        RowTile t;
        t.item = (int)k;
        t.row0 = r0;
        t.nrows = std::min(rows, heights[k] - r0);
        tiles.push_back(t);
      }
    }
  }

  int mci_label_slices(mci_ctx * ctx, int32_t n, const mci_slice_desc * descs, int fully, int32_t * n_components)
  {
    if (!ctx || !ctx->dIn || n < 0 || (n > 0 && !descs))
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->slices.assign(descs, descs + n);
    ctx->sliceDev.resize(n);
    ctx->ncomp.assign(n, 0);
    ctx->compBase.assign(n, 0);
    ctx->totalComp = 0;
    ctx->statsOnHost = false;
    ctx->hStats.clear();
    if (n == 0)
      return MCI_OK;
    const long long X = ctx->dims[0], Y = ctx->dims[1];
    u64 pix = 0, bitWords = 0, runTotal = 0;
    long long rows = 0;
    int maxRows = 0;
    std::vector<int> ws(n), hs(n);
    u64 compBound = 0;
    for (int k = 0; k < n; ++k)
    {
      const mci_slice_desc & d = descs[k];
      if (d.axis < 0 || d.axis > 2 || d.w < 1 || d.h < 1 || d.u0 < 0 || d.v0 < 0 || d.slice < 0 || d.slice >= ctx->dims[d.axis])
        return fail(ctx, MCI_ERR_ARG, "bad slice descriptor");
      const int d0 = d.axis == 0 ? 1 : 0, d1 = d.axis == 2 ? 1 : 2;
      if (d.u0 + d.w > ctx->dims[d0] || d.v0 + d.h > ctx->dims[d1])
        return fail(ctx, MCI_ERR_ARG, "slice region outside the volume");
      if ((u64)d.w * (u64)d.h >= 0xFFFFFFFFull)
        return fail(ctx, MCI_ERR_ARG, "slice region too large");
      SliceDev & S = ctx->sliceDev[k];
      S.axis = d.axis;
      S.slice = d.slice;
      S.label = (int)d.label;
      S.u0 = d.u0;
      S.v0 = d.v0;
      S.w = d.w;
      S.h = d.h;
      S.pixOff = pix;
      S.bitOff = bitWords;
      S.runOff = runTotal;
      S.runCap = (int)std::min<u64>((u64)(d.w / 2 + 1) * (u64)d.h, 4u << 20);
      bitWords += (u64)((d.w + 31) / 32) * d.h;
      runTotal += (u64)S.runCap;
      maxRows = std::max(maxRows, d.h);
      if (d.axis == 2)
      {
        S.su = 1;
        S.sv = X;
        S.base = (u64)d.slice * X * Y + (u64)d.v0 * X + d.u0;
      }
      else if (d.axis == 1)
      {
        S.su = 1;
        S.sv = X * Y;
        S.base = (u64)d.slice * X + (u64)d.v0 * X * Y + d.u0;
      }
      else
      {
        S.su = X;
        S.sv = X * Y;
        S.base = (u64)d.slice + (u64)d.u0 * X + (u64)d.v0 * X * Y;
      }
      pix += (u64)d.w * d.h;
      rows += d.h;
      ws[k] = d.w;
      hs[k] = d.h;
      compBound += ((u64)d.w * d.h + 1) / 2;
    }
    std::vector<RowTile> tiles;
    make_tiles(ws, hs, tiles);
    Blob blob;
    size_t oS = blob.add(ctx->sliceDev), oT = blob.add(tiles);
    CK(ctx->dSliceBlob.ensure(blob.bytes.size()));
    CK(ctx->hStage.ensure(blob.bytes.size()));
    std::memcpy(ctx->hStage.p, blob.bytes.data(), blob.bytes.size());
    cudaStream_t s = ctx->L.stream;
    {
      TableCopy tc;
      tc.add(ctx->dSliceBlob.p, ctx->hStage.p, blob.bytes.size());
      tc.launch(ctx->L);
    }
    ctx->dSlices = (SliceDev *)((uint8_t *)ctx->dSliceBlob.p + oS);
    RowTile * dTiles = (RowTile *)((uint8_t *)ctx->dSliceBlob.p + oT);
    ctx->statCap = (int)std::min<u64>(compBound, 2u << 20);
    CK(ctx->dBits.ensure((bitWords + 64) * 4));
    CK(ctx->dWordRun.ensure((bitWords + 64) * 4));
    CK(ctx->dRunComp.ensure((runTotal + 64) * 4));
    CK(ctx->dRunScratch.ensure((runTotal * 3 + 64) * 4));
    CK(ctx->dConn.ensure(pix * 2));
    CK(ctx->dNcomp.ensure(size_t(n) * 4));
    CK(ctx->dCompBase.ensure(size_t(n) * 4));
    CK(ctx->dTotal.ensure(4));
    CK(ctx->dStats.ensure(size_t(ctx->statCap) * sizeof(CompStat)));
    CK(cudaMemsetAsync(ctx->dTotal.p, 0, 4, s));
    launch_cc(ctx->L, ctx->ptype, ctx->dIn, ctx->dSlices, n, dTiles, (int)tiles.size(), (uint32_t *)ctx->dBits.p,
              (uint32_t *)ctx->dWordRun.p, (uint32_t *)ctx->dRunComp.p, (uint32_t *)ctx->dRunScratch.p, (uint16_t *)ctx->dConn.p,
              (int *)ctx->dNcomp.p, (int *)ctx->dCompBase.p, (int *)ctx->dTotal.p, (CompStat *)ctx->dStats.p, ctx->statCap, fully,
              maxRows, ctx->dErr);
    if (ctx->deferSync)
      return MCI_OK;
    CK(cudaMemcpyAsync(ctx->compBase.data(), ctx->dCompBase.p, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(ctx->ncomp.data(), ctx->dNcomp.p, size_t(n) * 4, cudaMemcpyDeviceToHost, s));
    int rc = check_device_error(ctx);
    if (rc)
      return rc;
    int acc = 0;
    for (int k = 0; k < n; ++k)
      acc += ctx->ncomp[k]; // compBase[] was filled by the device (slots are handed out in completion order)
    ctx->totalComp = acc;
    if (n_components)
      std::memcpy(n_components, ctx->ncomp.data(), size_t(n) * 4);
    return MCI_OK;
  }

  static int fetch_stats(mci_ctx * ctx)
  {
    if (ctx->statsOnHost)
      return MCI_OK;
    ctx->hStats.resize(ctx->totalComp);
    if (ctx->totalComp)
    {
      CK(cudaMemcpyAsync(ctx->hStats.data(), ctx->dStats.p, size_t(ctx->totalComp) * sizeof(CompStat), cudaMemcpyDeviceToHost,
                         ctx->L.stream));
      CK(cudaStreamSynchronize(ctx->L.stream));
    }
    ctx->statsOnHost = true;
    return MCI_OK;
  }

  int mci_get_components(mci_ctx * ctx, mci_component_info * out)
  {
    if (!ctx || !out)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    int rc = fetch_stats(ctx);
    if (rc)
      return rc;
    int k = 0;
    for (size_t sl = 0; sl < ctx->ncomp.size(); ++sl)
      for (int q = 0; q < ctx->ncomp[sl]; ++q, ++k)
      {
        const CompStat & s = ctx->hStats[ctx->compBase[sl] + q];
        out[k].count = s.count;
        out[k].sum_u = s.sumU;
        out[k].sum_v = s.sumV;
        out[k].min_u = s.minU;
        out[k].min_v = s.minV;
        out[k].max_u = s.maxU;
        out[k].max_v = s.maxV;
      }
    return MCI_OK;
  }

  int mci_get_component_image(mci_ctx * ctx, int32_t k, uint16_t * ids)
  {
    if (!ctx || !ids || k < 0 || k >= (int)ctx->slices.size())
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    const SliceDev & S = ctx->sliceDev[k];
    CK(cudaMemcpyAsync(ids, (uint16_t *)ctx->dConn.p + S.pixOff, size_t(S.w) * S.h * 2, cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return MCI_OK;
  }

  // ---- component matching ---------------------------------------------------------------------
  int mci_match_pairs(mci_ctx * ctx, int32_t nseg, const mci_segment_desc * segs, int32_t * n_pairs)
  {
    if (!ctx || nseg < 0 || (nseg > 0 && !segs))
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->nPairs = 0;
    if (n_pairs)
      *n_pairs = 0;
    if (nseg == 0)
      return MCI_OK;
    std::vector<SegDev> sd(nseg);
    std::vector<int> ws(nseg), hs(nseg);
    u64 bound = 0;
    for (int k = 0; k < nseg; ++k)
    {
      int a = segs[k].slice_i, b = segs[k].slice_j;
      if (a < 0 || b < 0 || a >= (int)ctx->slices.size() || b >= (int)ctx->slices.size())
        return fail(ctx, MCI_ERR_ARG, "segment refers to an unknown slice");
      const SliceDev &A = ctx->sliceDev[a], &B = ctx->sliceDev[b];
      if (A.w != B.w || A.h != B.h || A.u0 != B.u0 || A.v0 != B.v0)
        return fail(ctx, MCI_ERR_ARG, "the two slices of a segment must share one region");
      sd[k].si = a;
      sd[k].sj = b;
      ws[k] = A.w;
      hs[k] = A.h;
      // distinct pairs of a segment: at most (ni+1)(nj+1); before the counts are known, a bound from the area
      bound += ctx->deferSync ? std::min<u64>((u64)A.w * A.h / 16 + 64, 1u << 20)
                              : std::min<u64>((u64)(ctx->ncomp[a] + 1) * (u64)(ctx->ncomp[b] + 1), (u64)A.w * A.h);
    }
    std::vector<RowTile> tiles;
    make_tiles(ws, hs, tiles);
    Blob blob;
    size_t oS = blob.add(sd), oT = blob.add(tiles);
    CK(ctx->dSegBlob.ensure(blob.bytes.size()));
    CK(ctx->hStage2.ensure(blob.bytes.size()));
    std::memcpy(ctx->hStage2.p, blob.bytes.data(), blob.bytes.size());
    cudaStream_t s = ctx->L.stream;
    {
      TableCopy tc;
      tc.add(ctx->dSegBlob.p, ctx->hStage2.p, blob.bytes.size());
      tc.launch(ctx->L);
    }
    u64 tsize = 1 << 14;
    while (tsize < 4 * bound && tsize < (1ull << 27))
      tsize <<= 1;
    ctx->capPairs = (int)std::min<u64>(bound + 16, tsize / 2);
    CK(ctx->dPairTable.ensure(tsize * 8));
    CK(ctx->dPairs.ensure(size_t(ctx->capPairs) * sizeof(mci_pair)));
    CK(ctx->dCounters.ensure(2 * sizeof(int)));
    CK(cudaMemsetAsync(ctx->dPairTable.p, 0xFF, tsize * 8, s));
    CK(cudaMemsetAsync(ctx->dCounters.p, 0, sizeof(int), s));
    launch_pairs(ctx->L, ctx->dSlices, (SegDev *)((uint8_t *)ctx->dSegBlob.p + oS), (RowTile *)((uint8_t *)ctx->dSegBlob.p + oT),
                 (int)tiles.size(), (uint16_t *)ctx->dConn.p, (u64 *)ctx->dPairTable.p, (uint32_t)(tsize - 1),
                 (mci_pair *)ctx->dPairs.p, ctx->capPairs, (int *)ctx->dCounters.p, ctx->dErr);
    ctx->resPairs = 0;
    if (ctx->deferSync)
      return MCI_OK;
    int h = 0;
    CK(cudaMemcpyAsync(&h, ctx->dCounters.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    int rc = check_device_error(ctx);
    if (rc)
      return rc;
    ctx->nPairs = h;
    if (n_pairs)
      *n_pairs = h;
    return MCI_OK;
  }

  // Connected components and component matching launched back to back, ONE host round trip: component counts,
  // the pair list and the component statistics come back together (speculatively sized; anything beyond the
  // speculation is fetched on demand by the getters).
  int mci_label_and_match(mci_ctx * ctx, int32_t n, const mci_slice_desc * descs, int fully, int32_t nseg,
                          const mci_segment_desc * segs, int32_t * n_components, int32_t * n_pairs)
  {
    if (!ctx)
      return MCI_ERR_ARG;
    if (n > kSpecNcomp)
    {
      int rc = mci_label_slices(ctx, n, descs, fully, n_components);
      return rc ? rc : mci_match_pairs(ctx, nseg, segs, n_pairs);
    }
    ctx->deferSync = true;
    int rc = mci_label_slices(ctx, n, descs, fully, nullptr);
    if (!rc)
      rc = mci_match_pairs(ctx, nseg, segs, nullptr);
    ctx->deferSync = false;
    if (rc)
      return rc;
    if (n_pairs)
      *n_pairs = 0;
    if (n == 0)
      return MCI_OK;
    cudaStream_t s = ctx->L.stream;
    CK(ctx->hRes.ensure(kResBytes));
    uint8_t * hr = (uint8_t *)ctx->hRes.p;
    int * hn = (int *)(hr + kResNcompOff);
    int * hc = (int *)hr + 4; // [4]: pair count
    const int specPairs = nseg > 0 ? std::min(kSpecPairs, ctx->capPairs) : 0;
    const int specStats = std::min(kSpecStats, ctx->statCap);
    TableCopy tc;
    tc.add(hn, ctx->dNcomp.p, size_t(n) * 4);
    tc.add(hn + kSpecNcomp, ctx->dCompBase.p, size_t(n) * 4);
    tc.add(hr + kResStatsOff, ctx->dStats.p, size_t(specStats) * sizeof(CompStat), (int *)ctx->dTotal.p, sizeof(CompStat));
    if (nseg > 0)
    {
      tc.add(hc, ctx->dCounters.p, sizeof(int));
      tc.add(hr + kResPairsOff, ctx->dPairs.p, size_t(specPairs) * sizeof(mci_pair), (int *)ctx->dCounters.p, sizeof(mci_pair));
    }
    else
      *hc = 0;
    rc = check_device_error(ctx, &tc); // one launch: the five tables and the error flag
    if (rc)
      return rc;
    int acc = 0;
    for (int k = 0; k < n; ++k)
    {
      ctx->ncomp[k] = hn[k];
      ctx->compBase[k] = hn[kSpecNcomp + k];
      acc += hn[k];
    }
    ctx->totalComp = acc;
    ctx->nPairs = *hc;
    ctx->resPairs = std::min(ctx->nPairs, specPairs);
    if (acc <= specStats)
    {
      ctx->hStats.resize(acc);
      if (acc)
        std::memcpy(ctx->hStats.data(), hr + kResStatsOff, size_t(acc) * sizeof(CompStat));
      ctx->statsOnHost = true;
    }
    if (n_components)
      std::memcpy(n_components, ctx->ncomp.data(), size_t(n) * 4);
    if (n_pairs)
      *n_pairs = ctx->nPairs;
    return MCI_OK;
  }

  int mci_get_pairs(mci_ctx * ctx, mci_pair * pairs)
  {
    if (!ctx || (!pairs && ctx->nPairs))
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->nPairs && ctx->resPairs >= ctx->nPairs)
      std::memcpy(pairs, (const uint8_t *)ctx->hRes.p + kResPairsOff, size_t(ctx->nPairs) * sizeof(mci_pair));
    else if (ctx->nPairs)
    {
      CK(cudaMemcpyAsync(pairs, ctx->dPairs.p, size_t(ctx->nPairs) * sizeof(mci_pair), cudaMemcpyDeviceToHost, ctx->L.stream));
      CK(cudaStreamSynchronize(ctx->L.stream));
    }
    return MCI_OK;
  }

  // ---- interpolation --------------------------------------------------------------------------
  static int depth_of_gap(int g)
  {
    int d = 1;
    while (g > 2)
    {
      g = (g + 1) / 2;
      if (g <= 1)
        break;
      ++d;
    }
    return d;
  }

  // Volume writes of the mid slices perpendicular to z, then of those perpendicular to y (two passes: a voxel can lie
  // on slices of both kinds, and a pass assumes it is the only writer): tables prepared here, kernels in mci_dt.cu.
  static int compose_axes(mci_ctx * ctx, const EmitRec * recs, const int * nrecsDev)
  {
    cudaStream_t s = ctx->L.stream;
    const size_t maxPlanes = (size_t)std::max(ctx->dims[1], ctx->dims[2]);
    CK(ctx->dCompose.ensure((7 * maxPlanes + 1 + (size_t)ctx->emitCap + 8) * sizeof(int)));
    int * tables = (int *)ctx->dCompose.p;
    for (int axis = 2; axis >= 1; --axis)
    {
      if (!ctx->axisJobs[axis])
        continue; // no slices perpendicular to this axis in the run
      const size_t nP = (size_t)ctx->dims[axis];
      if (!compose_tables_in_one_launch(ctx->emitCap)) // else k_compose_tables initialises them itself
      {
        CK(cudaMemsetAsync(tables, 0, (3 * nP + 1) * sizeof(int), s));         // count, start, fill
        CK(cudaMemsetAsync(tables + 3 * nP + 1, 0x7F, 4 * nP * sizeof(int), s)); // box rows {min x, min y, max x, max y}: minima ...
        // ... and maxima, which k_compose_count raises with atomicMax: they must start below any coordinate
        CK(cudaMemset2DAsync(tables + 3 * nP + 1 + 2, 4 * sizeof(int), 0x80, 2 * sizeof(int), nP, s));
      }
      launch_compose(ctx->L, ctx->ptype, axis, recs, nrecsDev, ctx->emitCap, (const uint32_t *)ctx->dArena.p + ctx->emitBase, tables,
                     ctx->dIn, ctx->dOut, ctx->dims);
    }
    return MCI_OK;
  }

  int mci_interpolate(mci_ctx * ctx, int32_t njobs, const mci_job * jobs, const int32_t * b_ids, int32_t n_b_ids,
                      const mci_params * prm)
  {
    if (!ctx || !prm || njobs < 0 || (njobs > 0 && !jobs))
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    ctx->nJobs = njobs;
    ctx->haveEmit = false;
    ctx->axisJobs[0] = ctx->axisJobs[1] = ctx->axisJobs[2] = 0;
    if (njobs == 0)
      return MCI_OK;
    int rc = fetch_stats(ctx);
    if (rc)
      return rc;
    const int nsl = (int)ctx->slices.size();
    cudaStream_t s = ctx->L.stream;

    std::vector<MaskRef> refs;
    std::vector<ExtractJob> extracts;
    std::vector<AlignJob> aligns(njobs);
    std::vector<SplitJob> splits;
    std::vector<RootJob> roots(njobs);
    // mask of a component, made at its first use: looked up by the component's dense index (compBase[slice] + id - 1),
    // not through a map -- this loop runs on the host between two device phases of every step
    std::vector<int> compMaskAt(ctx->hStats.size(), -1);
    std::vector<MaskRef> compMaskStore;
    refs.reserve(size_t(njobs) * 3 + 16);
    extracts.reserve(size_t(njobs) * 4 + 16);
    compMaskStore.reserve(size_t(njobs) * 2 + 16);
    u64 arenaWords = 0, scratchWords = 0;
    int maxSplitWords = 0;
    int nSlots = 0; // axis-0 slot table entries
    std::vector<Axis0Group> axis0Groups;
    const int alignSmemCap = 150 * 1024; // bitmap + queue; the kernel adds 48 KB for staging the shapes
    int alignSmem = 0;
    int nRootNodes = 0;
    int maxDepth = 0;
    std::vector<long long> levelCap;
    std::vector<unsigned long long> levelScratch, levelItems, levelHItems;
    u64 midWords = 0, maxNodeWords = 0;

    auto statOf = [&](int slice, int id) -> const CompStat & { return ctx->hStats[ctx->compBase[slice] + id - 1]; };
    auto compMask = [&](int slice, int id) -> MaskRef {
      int & at = compMaskAt[size_t(ctx->compBase[slice] + id - 1)];
      if (at >= 0)
        return compMaskStore[size_t(at)];
      const CompStat & c = statOf(slice, id);
      MaskRef m;
      m.x0 = c.minU;
      m.y0 = c.minV;
      m.w = c.maxU - c.minU + 1;
      m.h = c.maxV - c.minV + 1;
      m.pitch = (m.w + 31) >> 5;
      m.pad = 0;
      m.off = arenaWords;
      arenaWords += (u64)m.h * m.pitch;
      const int tileRows = std::max(1, 256 / m.pitch);
      for (int r0 = 0; r0 < m.h; r0 += tileRows)
      {
        ExtractJob e;
        e.slice = slice;
        e.id = id;
        e.row0 = r0;
        e.nrows = std::min(tileRows, m.h - r0);
        e.m = m;
        extracts.push_back(e);
      }
      at = (int)compMaskStore.size();
      compMaskStore.push_back(m);
      return m;
    };

    for (int k = 0; k < njobs; ++k)
    {
      const mci_job & jb = jobs[k];
      if (jb.kind < 0 || jb.kind > 2 || jb.slice_a < 0 || jb.slice_a >= nsl || jb.a_id < 1 || jb.a_id > ctx->ncomp[jb.slice_a])
        return fail(ctx, MCI_ERR_ARG, "bad job (a side)");
      const SliceDev & SA = ctx->sliceDev[jb.slice_a];
      const int gap = std::abs(jb.pos_a - jb.pos_b);
      if (gap < 2)
        return fail(ctx, MCI_ERR_ARG, "job between adjacent slices");
      AlignJob & A = aligns[k];
      RootJob & R = roots[k];
      std::memset(&A, 0, sizeof(A));
      std::memset(&R, 0, sizeof(R));
      A.kind = R.kind = jb.kind;
      A.out = k;
      R.job = k;
      R.posA = jb.pos_a;
      R.posB = jb.pos_b;
      R.axis = SA.axis;
      ctx->axisJobs[SA.axis]++;
      R.label = SA.label;
      R.rootOff = nRootNodes;
      R.slot0 = -1;
      const CompStat & ca = statOf(jb.slice_a, jb.a_id);
      const long long acu = (long long)((unsigned long long)ca.sumU / (unsigned long long)ca.count);
      const long long acv = (long long)((unsigned long long)ca.sumV / (unsigned long long)ca.count);
      MaskRef ma = compMask(jb.slice_a, jb.a_id);
      A.aRef = R.aRef = (int)refs.size();
      refs.push_back(ma);
      int rbU0, rbV0, rbW, rbH; // region of the "j" image handed to Align
      long long wbSum = 0, hbMax = 0;
      int nRootsHere = 1;
      if (jb.kind == JOB_EXTRAPOLATE)
      {
        A.nB = 0;
        A.indx = A.indy = 0; // both centroids coincide (X:216,240)
        A.phx = R.phx = (int)acu;
        A.phy = R.phy = (int)acv;
        rbU0 = (int)acu - 1;
        rbV0 = (int)acv - 1;
        rbW = rbH = 3;
        MaskRef ph;
        ph.x0 = ph.y0 = 0;
        ph.w = ph.h = 1;
        ph.pitch = 1;
        ph.pad = 0;
        ph.off = arenaWords;
        arenaWords += 1;
        R.phRef = (int)refs.size();
        refs.push_back(ph);
        wbSum = 1;
        hbMax = 1;
      }
      else
      {
        if (jb.slice_b < 0 || jb.slice_b >= nsl || jb.n_b < 1 || jb.b_ids_offset < 0 || jb.b_ids_offset + jb.n_b > n_b_ids ||
            (jb.kind == JOB_ONE_TO_ONE && jb.n_b != 1))
          return fail(ctx, MCI_ERR_ARG, "bad job (b side)");
        const SliceDev & SB = ctx->sliceDev[jb.slice_b];
        rbU0 = SB.u0;
        rbV0 = SB.v0;
        rbW = SB.w;
        rbH = SB.h;
        A.nB = R.nB = jb.n_b;
        A.bRef = R.bRef = (int)refs.size();
        unsigned long long su = 0, sv = 0, cnt = 0;
        for (int q = 0; q < jb.n_b; ++q)
        {
          int id = b_ids[jb.b_ids_offset + q];
          if (id < 1 || id > ctx->ncomp[jb.slice_b])
            return fail(ctx, MCI_ERR_ARG, "bad job (b id)");
          const CompStat & cb = statOf(jb.slice_b, id);
          su += (unsigned long long)cb.sumU;
          sv += (unsigned long long)cb.sumV;
          cnt += (unsigned long long)cb.count;
          MaskRef mb = compMask(jb.slice_b, id);
          refs.push_back(mb);
          wbSum = std::max<long long>(wbSum, mb.w);
          hbMax = std::max<long long>(hbMax, mb.h);
        }
        A.indx = (int)((long long)(su / cnt) - acu); // X:1147-1154
        A.indy = (int)((long long)(sv / cnt) - acv);
        if (jb.kind == JOB_ONE_TO_N)
        {
          SplitJob sp;
          sp.job = k;
          sp.aRef = A.aRef;
          sp.nB = jb.n_b;
          sp.bRef = A.bRef;
          sp.outRef = (int)refs.size();
          sp.pad = 0;
          sp.scratchOff = scratchWords;
          scratchWords += (u64)jb.n_b * ma.h * ma.pitch;
          maxSplitWords = (int)std::min<u64>(std::max<u64>((u64)maxSplitWords, (u64)ma.h * ma.pitch), 1u << 20);
          R.splitRef = sp.outRef;
          for (int q = 0; q < jb.n_b; ++q)
          {
            MaskRef mo = ma;
            mo.off = arenaWords;
            arenaWords += (u64)ma.h * ma.pitch;
            refs.push_back(mo);
          }
          splits.push_back(sp);
          nRootsHere = jb.n_b;
        }
      }
      // search region of translations (X:1156-1164)
      A.sx0 = rbU0 - SA.u0 - SA.w + 1;
      A.sy0 = rbV0 - SA.v0 - SA.h + 1;
      A.sw = SA.w + rbW - 1;
      A.sh = SA.h + rbH - 1;
      const unsigned long long nS = (unsigned long long)A.sw * (unsigned long long)A.sh;
      A.minIter = (long long)std::min<unsigned long long>((unsigned long long)prm->min_align_iters, nS);
      A.maxIter = (long long)std::max<unsigned long long>((unsigned long long)prm->max_align_iters,
                                                        (unsigned long long)std::sqrt((double)nS));
      // ring capacity: with the heuristic on and a non-zero score at t=0 (guaranteed for shapes paired by pixel
      // overlap) at most maxIter + 4 translations are ever pushed; otherwise two full BFS rings can be queued
      long long need = (prm->heuristic_alignment && jb.kind != JOB_EXTRAPOLATE)
                         ? A.maxIter + 16
                         : std::max<long long>(A.maxIter + 16, 8ll * (A.sw + A.sh)) + 64;
      int qcap = 64;
      while (qcap < need)
        qcap <<= 1;
      A.qcap = qcap;
      const u64 bwords = (u64)((A.sw + 31) >> 5) * (u64)A.sh;
      const u64 smemNeed = (bwords + 3ull * qcap) * 4;
      if (smemNeed <= (u64)alignSmemCap)
      {
        A.useSmem = 1;
        alignSmem = std::max(alignSmem, (int)smemNeed);
      }
      else
      {
        A.useSmem = 0;
        A.bitmapOff = scratchWords;
        scratchWords += bwords;
        A.queueOff = scratchWords;
        scratchWords += 3ull * qcap;
      }
      // trees perpendicular to x register their mid slices by position (k_apply_axis0)
      if (SA.axis == 0)
      {
        R.slot0 = nSlots;
        for (int q = 0; q < nRootsHere; ++q)
        {
          for (int c0 = 0; c0 < gap - 1; c0 += 32)
          {
            Axis0Group g;
            g.slot0 = nSlots + c0;
            g.n = std::min(32, gap - 1 - c0);
            axis0Groups.push_back(g);
          }
          nSlots += gap - 1;
        }
      }
      // node bookkeeping
      nRootNodes += nRootsHere;
      int depth = depth_of_gap(gap);
      maxDepth = std::max(maxDepth, depth);
      if ((int)levelCap.size() < depth)
      {
        levelCap.resize(depth, 0);
        levelScratch.resize(depth, 0);
        levelItems.resize(depth, 0);
        levelHItems.resize(depth, 0);
      }
      const int d0 = SA.axis == 0 ? 1 : 0, d1 = SA.axis == 2 ? 1 : 2;
      long long bw = std::min<long long>(ctx->dims[d0], 2 * (ma.w + wbSum));
      long long bh = std::min<long long>(ctx->dims[d1], 2 * (ma.h + hbMax));
      u64 nodeWords = (u64)((bw + 31) / 32 + 1) * (u64)bh;
      maxNodeWords = std::max(maxNodeWords, nodeWords);
      midWords += (u64)nRootsHere * (u64)(gap - 1) * nodeWords;
      for (int d = 0; d < depth; ++d)
      {
        const long long nn = (long long)nRootsHere * std::min<long long>(1ll << std::min(d, 30), gap - 1);
        levelCap[d] += nn;
        levelScratch[d] += (u64)nn * (3 * nodeWords + (u64)bh + (u64)bh / 8 + 16);
        levelItems[d] += (u64)nn * (nodeWords / 256 + 1);
        levelHItems[d] += (u64)nn * (nodeWords / MCI_DT_HCHUNK + 1);
      }
    }

    // alignment jobs of the 1-to-N kind go first (their splits overlap the rest of the alignment)
    std::stable_partition(aligns.begin(), aligns.end(), [](const AlignJob & a) { return a.kind == JOB_ONE_TO_N; });
    const int nSplitJobs = (int)splits.size();

    // ---- device memory
    const u64 arenaCap = arenaWords + midWords + 64;
    Blob blob;
    const size_t oRefs = blob.add(refs), oExtract = blob.add(extracts), oAlign = blob.add(aligns), oSplit = blob.add(splits),
                 oRoots = blob.add(roots);
    std::vector<int> lc(maxDepth + 2, 0);
    lc[0] = nRootNodes;
    const std::vector<unsigned long long> top(1, arenaWords);
    const size_t oLc = blob.add(lc), oTop = blob.add(top), oGroups = blob.add(axis0Groups);
    CK(ctx->dJobBlob.ensure(blob.bytes.size()));
    CK(ctx->hStage.ensure(blob.bytes.size()));
    CK(ctx->dLevelCounts.ensure(size_t(maxDepth + 2) * sizeof(int)));
    CK(ctx->dArenaTop.ensure(8));
    std::memcpy(ctx->hStage.p, blob.bytes.data(), blob.bytes.size());
    {
      TableCopy tc;
      tc.add(ctx->dJobBlob.p, ctx->hStage.p, blob.bytes.size());
      tc.add(ctx->dLevelCounts.p, (uint8_t *)ctx->hStage.p + oLc, lc.size() * sizeof(int));
      tc.add(ctx->dArenaTop.p, (uint8_t *)ctx->hStage.p + oTop, 8);
      tc.launch(ctx->L);
    }
    uint8_t * jb = (uint8_t *)ctx->dJobBlob.p;
    MaskRef * dRefs = (MaskRef *)(jb + oRefs);
    CK(ctx->dArena.ensure(arenaCap * 4));
    CK(ctx->dScratch.ensure((scratchWords + 64) * 4));
    CK(ctx->dTrans.ensure(size_t(njobs) * (sizeof(int2) + sizeof(int4)) + 64));
    long long nodeTotal = 0;
    for (long long c : levelCap)
      nodeTotal += c;
    CK(ctx->dNodes.ensure(size_t(nodeTotal + 1) * sizeof(Node)));
    CK(ctx->dEmit.ensure(16 + size_t(nodeTotal + 1) * sizeof(EmitRec) + size_t(nSlots + 1) * sizeof(int)));
    CK(cudaMemsetAsync(ctx->dEmit.p, 0, 16, s));
    ctx->emitCap = (int)nodeTotal;
    ctx->emitBase = arenaWords;
    ctx->haveEmit = true;
    EmitRec * dEmitRecs = (EmitRec *)((uint8_t *)ctx->dEmit.p + 16);
    int * dEmitCount = (int *)ctx->dEmit.p;
    int * dSlots = emit_slot_table(dEmitRecs, ctx->emitCap);
    if (nSlots)
      CK(cudaMemsetAsync(dSlots, 0xFF, size_t(nSlots) * sizeof(int), s)); // -1: no mid slice at that position

    // ---- launches
    launch_extract(ctx->L, ctx->dSlices, (ExtractJob *)(jb + oExtract), (int)extracts.size(), (uint16_t *)ctx->dConn.p,
                   (uint32_t *)ctx->dArena.p);
    // The 1-to-N jobs (alignment, then the split) take a second stream: they are few, and their chain is longer than
    // the alignment of all the other jobs, which runs beside it on the main stream
    int4 * dDbg = (int4 *)((uint8_t *)ctx->dTrans.p + ((size_t(njobs) * sizeof(int2) + 15) & ~size_t(15)));
    if (!splits.empty())
    {
      CK(cudaEventRecord(ctx->evAlignedSplits, s)); // masks extracted
      CK(cudaStreamWaitEvent(ctx->stream2, ctx->evAlignedSplits, 0));
      ctx->L.stream = ctx->stream2;
      launch_align(ctx->L, (AlignJob *)(jb + oAlign), nSplitJobs, dRefs, (uint32_t *)ctx->dArena.p, (uint32_t *)ctx->dScratch.p,
                   prm->heuristic_alignment, (int2 *)ctx->dTrans.p, dDbg, alignSmem, ctx->dErr);
      launch_split(ctx->L, (SplitJob *)(jb + oSplit), (int)splits.size(), dRefs, (uint32_t *)ctx->dArena.p,
                   (uint32_t *)ctx->dScratch.p, (int2 *)ctx->dTrans.p, prm->use_ball, maxSplitWords, ctx->dErr);
      ctx->L.stream = s;
      CK(cudaEventRecord(ctx->evSplitDone, ctx->stream2));
    }
    launch_align(ctx->L, (AlignJob *)(jb + oAlign) + nSplitJobs, njobs - nSplitJobs, dRefs, (uint32_t *)ctx->dArena.p,
                 (uint32_t *)ctx->dScratch.p, prm->heuristic_alignment, (int2 *)ctx->dTrans.p, dDbg, alignSmem, ctx->dErr);
    if (!splits.empty())
      CK(cudaStreamWaitEvent(s, ctx->evSplitDone, 0));
    Node * dNodes = (Node *)ctx->dNodes.p;
    launch_make_roots(ctx->L, (RootJob *)(jb + oRoots), njobs, dRefs, (uint32_t *)ctx->dArena.p, (int2 *)ctx->dTrans.p, dNodes);
    long long off = 0;
    int * dCounts = (int *)ctx->dLevelCounts.p;
    if (prm->use_distance_transform)
    {
      // distance-transform median: one persistent kernel over a device-side node queue + compose passes (mci_dt.cu);
      // MCI_DT_LEVELS=1 selects the round-1 form, six grid-wide kernels per level
      long long maxLevelNodes = 0;
      u64 scratchCap = 0, itemCap = 0, hitemCap = 0;
      for (int d = 0; d < maxDepth; ++d)
      {
        maxLevelNodes = std::max(maxLevelNodes, levelCap[d]);
        scratchCap = std::max<u64>(scratchCap, levelScratch[d]);
        itemCap = std::max<u64>(itemCap, levelItems[d]);
        hitemCap = std::max<u64>(hitemCap, levelHItems[d]);
      }
      const int histSide = ctx->ptype == MCI_U8 ? 256 : (ctx->ptype == MCI_U16 ? 6560 : 3296);
      const int histStride = 2 * histSide + histSide / 32; // counts of both sides + bitmap of touched entries
      static const bool useLevels = std::getenv("MCI_DT_LEVELS") != nullptr; // the level-synchronous path of round 1 (A/B runs)
      const int treeGridMax = ctx->L.sms * 2; // two 512-thread CTAs per SM (1024 x 1: C3 0.37 vs 0.41 ms, but C5 6.5 vs 5.0 ms)
      const int nbig = useLevels ? 128 : std::max(128, treeGridMax); // tree path: a CTA keeps one wrapped-bin table
      const size_t histBytes = useLevels ? size_t(maxLevelNodes) * histStride * 4 : 0;
      const size_t bigBytes = size_t(nbig) * (2 * 65536 + 2048) * 4;
      if (useLevels)
      {
        if (ctx->dHist.cap < histBytes)
          ctx->histClean = false;
        CK(ctx->dHist.ensure(histBytes));
        if (!ctx->histClean)
        {
          CK(cudaMemsetAsync(ctx->dHist.p, 0, ctx->dHist.cap, s));
          ctx->histClean = true;
        }
      }
      if (ctx->dDtBig.cap < bigBytes)
        ctx->dtBigClean = false;
      CK(ctx->dDtBig.ensure(bigBytes));
      if (!ctx->dtBigClean)
      {
        CK(cudaMemsetAsync(ctx->dDtBig.p, 0, bigBytes, s));
        ctx->dtBigClean = true;
      }
      if (!useLevels)
      {
        // one CTA per recursion tree (k_dt_tree): a node's rows live in shared memory when they fit, else in the
        // CTA's own slice of global scratch
        const u64 perCta = 4 * maxNodeWords + 64;
        // few nodes: one 1024-thread CTA per SM -- every phase of a node twice as wide, the launch lasts as long as the
        // chains of its biggest nodes; many nodes: two 512-thread CTAs per SM.  Measured k_dt_tree, 1024 / 512 threads:
        // 940 nodes (C3) 0.24 / 0.29 ms; 2 300 nodes (one of eight C5 slabs) 0.79 / 1.17; 5 000 nodes (one of four)
        // 1.02 / 0.96; 9 000 nodes 1.75 / 1.64; 17 900 nodes (C5) 3.29 / 2.56
        static const char * forceWidth = std::getenv("MCI_TREE_THREADS"); // A/B runs: 512 or 1024
        const int treeThreads = forceWidth ? (std::atoi(forceWidth) >= 1024 ? 1024 : 512) : (nodeTotal <= 28ll * ctx->L.sms ? 1024 : 512);
        int grid = (int)std::max<long long>(1, std::min<long long>(nodeTotal, treeThreads == 1024 ? ctx->L.sms : treeGridMax));
        while (grid > 1 && (u64)grid * perCta * 4 > (u64(2) << 30))
          grid = (grid + 1) / 2;
        CK(ctx->dLevelScratch.ensure(((u64)grid * perCta + 64) * 4));
        // counters (head, wrapped-bin tables in use, tail, pushed, done) followed by one "published" flag per child slot
        const long long childSlots = std::max<long long>(nodeTotal - nRootNodes, 0) + 1;
        CK(ctx->dDtCounters.ensure(size_t(8 + childSlots) * sizeof(int)));
        CK(cudaMemsetAsync(ctx->dDtCounters.p, 0, size_t(8 + childSlots) * sizeof(int), s));
        TreeParams P;
        P.roots = dNodes;
        P.rootCount = dCounts;
        P.rootNext = (int *)ctx->dDtCounters.p;
        P.flags = (int *)ctx->dDtCounters.p + 8;
        P.queueCap = (int)childSlots;
        P.arena = (uint32_t *)ctx->dArena.p;
        P.arenaTop = (unsigned long long *)ctx->dArenaTop.p;
        P.arenaCap = arenaCap;
        P.scratch = (uint32_t *)ctx->dLevelScratch.p;
        P.scratchPerCta = perCta;
        P.histSide = histSide;
        P.bigHist = (unsigned *)ctx->dDtBig.p;
        P.bigCount = (int *)ctx->dDtCounters.p + 1;
        P.nbig = nbig;
        P.emit = dEmitRecs;
        P.emitCount = dEmitCount;
        P.emitCap = ctx->emitCap;
        P.emitBase = ctx->emitBase;
        P.in = ctx->dIn;
        P.out = ctx->dOut;
        P.VX = ctx->dims[0];
        P.VY = ctx->dims[1];
        P.VZ = ctx->dims[2];
        P.err = ctx->dErr;
        // mid slices perpendicular to z are written by one compose pass after the recursion (one thread per voxel, no
        // atomics) instead of compare-and-swap groups inside it; MCI_NO_COMPOSE=1 keeps the writes in the nodes (A/B)
        static const bool useCompose = std::getenv("MCI_NO_COMPOSE") == nullptr;
        P.composeAxis2 = useCompose ? 1 : 0;
        if (std::getenv("MCI_TREE_PROFILE"))
          CK(cudaMemsetAsync(P.scratch + (size_t)grid * perCta, 0, 64, s));
        launch_dt_tree(ctx->L, ctx->ptype, P, nRootNodes, grid, treeThreads);
        if (std::getenv("MCI_TREE_PROFILE"))
        {
          // cycles per phase summed over the CTAs (only meaningful in a -DMCI_TREE_PROFILE build)
          unsigned long long ph[7];
          CK(cudaMemcpyAsync(ph, P.scratch + (size_t)grid * perCta, sizeof(ph), cudaMemcpyDeviceToHost, s));
          CK(cudaStreamSynchronize(s));
          std::fprintf(stderr, "k_dt_tree phases (cycles / CTA, grid %d): prep %llu hist %llu scan %llu median %llu alloc %llu emit %llu other %llu\n",
                       grid, ph[0] / grid, ph[1] / grid, ph[2] / grid, ph[3] / grid, ph[4] / grid, ph[5] / grid, ph[6] / grid);
        }
      }
      else
      {
        CK(ctx->dWork.ensure(size_t(maxLevelNodes) * sizeof(NodeWork)));
        CK(ctx->dItems.ensure(size_t(2 * itemCap + hitemCap + 4) * sizeof(DtItem)));
        CK(ctx->dLevelScratch.ensure((scratchCap + 64) * 4));
        // per level: [0] items, [1] write items, [2] wrapped-bin tables in use; and a 64-bit scratch top
        CK(ctx->dDtCounters.ensure(size_t(maxDepth) * (4 * sizeof(int) + 8)));
        CK(cudaMemsetAsync(ctx->dDtCounters.p, 0, size_t(maxDepth) * (4 * sizeof(int) + 8), s));
        unsigned long long * tops = (unsigned long long *)ctx->dDtCounters.p;
        int * ctrs = (int *)(tops + maxDepth);
        for (int d = 0; d < maxDepth; ++d)
        {
          DtLevel lv;
          lv.nodes = dNodes + off;
          lv.count = dCounts + d;
          lv.maxNodes = (int)levelCap[d];
          lv.work = (NodeWork *)ctx->dWork.p;
          lv.arena = (uint32_t *)ctx->dArena.p;
          lv.arenaTop = (unsigned long long *)ctx->dArenaTop.p;
          lv.arenaCap = arenaCap;
          lv.scratch = (uint32_t *)ctx->dLevelScratch.p;
          lv.scratchTop = tops + d;
          lv.scratchCap = scratchCap;
          lv.items = (DtItem *)ctx->dItems.p;
          lv.itemCount = ctrs + 4 * d;
          lv.witems = (DtItem *)ctx->dItems.p + itemCap + 1;
          lv.witemCount = ctrs + 4 * d + 1;
          lv.itemCap = (int)itemCap;
          lv.hitems = (DtItem *)ctx->dItems.p + 2 * itemCap + 2;
          lv.hitemCount = ctrs + 4 * d + 3;
          lv.hitemCap = (int)hitemCap;
          lv.hist = (unsigned *)ctx->dHist.p;
          lv.histStride = histStride;
          lv.histSide = histSide;
          lv.bigHist = (unsigned *)ctx->dDtBig.p;
          lv.bigCount = ctrs + 4 * d + 2;
          lv.nbig = nbig;
          lv.next = dNodes + off + levelCap[d];
          lv.nextCount = dCounts + d + 1;
          lv.nextCap = d + 1 < maxDepth ? (int)levelCap[d + 1] : 0;
          lv.emit = dEmitRecs;
          lv.emitCount = dEmitCount;
          lv.emitCap = ctx->emitCap;
          lv.emitBase = ctx->emitBase;
          lv.in = ctx->dIn;
          lv.out = ctx->dOut;
          for (int a = 0; a < 3; ++a)
            lv.dims[a] = ctx->dims[a];
          lv.err = ctx->dErr;
          launch_dt_level(ctx->L, ctx->ptype, lv);
          off += levelCap[d];
        }
      }
      if (!useLevels && std::getenv("MCI_NO_COMPOSE") == nullptr)
      {
        int rc2 = compose_axes(ctx, dEmitRecs, dEmitCount);
        if (rc2)
          return rc2;
      }
      // the trees perpendicular to x go x-run by x-run
      launch_apply_axis0(ctx->L, ctx->ptype, (Axis0Group *)(jb + oGroups), (int)axis0Groups.size(), dEmitRecs, dSlots,
                         (uint32_t *)ctx->dArena.p + ctx->emitBase, ctx->dIn, ctx->dOut, ctx->dims);
    }
    else
    {
      const int nodeSmem = 100 * 1024; // two CTAs per SM: shapes up to ~4 200 words (e.g. 360 x 360 pixels) stay on chip
      const int nodeGrid = ctx->L.sms * 2;
      if (ctx->bigHistGrid < nodeGrid)
      {
        CK(ctx->dBigHist.ensure(size_t(nodeGrid) * 2 * 65536 * 4));
        CK(cudaMemsetAsync(ctx->dBigHist.p, 0, size_t(nodeGrid) * 2 * 65536 * 4, s));
        ctx->bigHistGrid = nodeGrid;
      }
      const u64 smemWords = (u64)nodeSmem / 4;
      const u64 needWords = maxNodeWords * 6 + 6 * (maxNodeWords / 32 + 1);
      u64 slabWords = 0;
      if (needWords > smemWords)
      {
        slabWords = needWords;
        CK(ctx->dSlab.ensure(size_t(nodeGrid) * slabWords * 4));
      }
      // arrival-step maps of the two dilation sequences, one pair per resident CTA (capped: larger shapes re-run)
      const u64 arrPixels = std::min<u64>(maxNodeWords * 32, 4u << 20);
      CK(ctx->dArrival.ensure(size_t(nodeGrid) * 2 * arrPixels * sizeof(unsigned short)));
      for (int d = 0; d < maxDepth; ++d)
      {
        Node * cur = dNodes + off;
        Node * nxt = dNodes + off + levelCap[d];
        int nextCap = d + 1 < maxDepth ? (int)levelCap[d + 1] : 0;
        launch_nodes(ctx->L, ctx->ptype, cur, dCounts + d, (int)levelCap[d], nxt, dCounts + d + 1, nextCap,
                     (uint32_t *)ctx->dArena.p, (unsigned long long *)ctx->dArenaTop.p, arenaCap, ctx->dIn, ctx->dOut, ctx->dims,
                     0, prm->use_ball, (unsigned *)ctx->dBigHist.p, (uint32_t *)ctx->dSlab.p, slabWords,
                     (unsigned short *)ctx->dArrival.p, arrPixels, nodeGrid, nodeSmem,
                     dEmitRecs, dEmitCount, ctx->emitCap, ctx->emitBase, ctx->dErr);
        off += levelCap[d];
      }
      // the node kernel only stores the mid shapes; their volume writes (X:743-764) happen here: slices perpendicular to z
      // and y through the compose passes (MCI_NO_COMPOSE=1: one grid-wide pass of compare-and-swap groups, as in round 1)
      if (std::getenv("MCI_NO_COMPOSE") == nullptr)
      {
        int rc2 = compose_axes(ctx, dEmitRecs, dEmitCount);
        if (rc2)
          return rc2;
      }
      else
        launch_apply_masks(ctx->L, ctx->ptype, dEmitRecs, ctx->emitCap, dEmitCount, (uint32_t *)ctx->dArena.p + ctx->emitBase, ctx->dIn,
                           ctx->dOut, ctx->dims, /*skipAxis0=*/1);
      launch_apply_axis0(ctx->L, ctx->ptype, (Axis0Group *)(jb + oGroups), (int)axis0Groups.size(), dEmitRecs, dSlots,
                         (uint32_t *)ctx->dArena.p + ctx->emitBase, ctx->dIn, ctx->dOut, ctx->dims);
    }
    CK(cudaGetLastError());
    return MCI_OK;
  }

  // per job: pops, waves, clock cycles, kind*1000 + n_b (performance diagnostics)
  int mci_get_align_stats(mci_ctx * ctx, int32_t * out /* [4*n_jobs] */)
  {
    if (!ctx || !out)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->nJobs)
    {
      CK(cudaMemcpyAsync(out, (uint8_t *)ctx->dTrans.p + ((size_t(ctx->nJobs) * sizeof(int2) + 15) & ~size_t(15)),
                         size_t(ctx->nJobs) * 16, cudaMemcpyDeviceToHost, ctx->L.stream));
      CK(cudaStreamSynchronize(ctx->L.stream));
    }
    return MCI_OK;
  }

  // ---- multi-GPU combine by mid-slice shapes ---------------------------------------------------------
  int mci_get_emitted(mci_ctx * ctx, void ** recs_dev, int64_t * n_recs, void ** words_dev, int64_t * n_words)
  {
    if (!ctx || !recs_dev || !n_recs || !words_dev || !n_words)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    *recs_dev = nullptr;
    *words_dev = nullptr;
    *n_recs = 0;
    *n_words = 0;
    if (!ctx->haveEmit)
      return MCI_OK; // the last run had no interpolation job
    int cnt = 0;
    unsigned long long top = 0;
    CK(cudaMemcpyAsync(&cnt, ctx->dEmit.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->L.stream));
    CK(cudaMemcpyAsync(&top, ctx->dArenaTop.p, 8, cudaMemcpyDeviceToHost, ctx->L.stream));
    int rc = check_device_error(ctx);
    if (rc)
      return rc;
    *recs_dev = (uint8_t *)ctx->dEmit.p + 16;
    *n_recs = std::min(cnt, ctx->emitCap);
    *words_dev = (uint32_t *)ctx->dArena.p + ctx->emitBase;
    *n_words = (int64_t)(top - ctx->emitBase);
    return MCI_OK;
  }

  int mci_copy_emitted(mci_ctx * ctx, void * dst_dev, int64_t words_offset_bytes)
  {
    if (!ctx || !dst_dev || words_offset_bytes < 0)
      return MCI_ERR_ARG;
    void *recs = nullptr, *words = nullptr;
    int64_t nr = 0, nw = 0;
    int rc = mci_get_emitted(ctx, &recs, &nr, &words, &nw);
    if (rc)
      return rc;
    if ((int64_t)(nr * sizeof(EmitRec)) > words_offset_bytes)
      return fail(ctx, MCI_ERR_ARG, "record area too small");
    if (nr)
      CK(cudaMemcpyAsync(dst_dev, recs, size_t(nr) * sizeof(EmitRec), cudaMemcpyDeviceToDevice, ctx->L.stream));
    if (nw)
      CK(cudaMemcpyAsync((uint8_t *)dst_dev + words_offset_bytes, words, size_t(nw) * 4, cudaMemcpyDeviceToDevice, ctx->L.stream));
    CK(cudaStreamSynchronize(ctx->L.stream));
    return MCI_OK;
  }

  int mci_apply_masks(mci_ctx * ctx, const void * recs_dev, int64_t n_recs, const void * words_dev, int64_t n_words)
  {
    if (!ctx || n_recs < 0 || (n_recs > 0 && (!recs_dev || !words_dev)) || !ctx->dIn || !ctx->dOut)
      return MCI_ERR_ARG;
    (void)n_words;
    CK(cudaSetDevice(ctx->device));
    launch_apply_masks(ctx->L, ctx->ptype, (const EmitRec *)recs_dev, (int)n_recs, nullptr, (const uint32_t *)words_dev, ctx->dIn,
                       ctx->dOut, ctx->dims, /*skipAxis0=*/0); // another GPU's list: no slot table, slice by slice
    return MCI_OK;
  }

  int mci_get_translations(mci_ctx * ctx, int32_t * t)
  {
    if (!ctx || !t)
      return MCI_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    if (ctx->nJobs)
    {
      CK(cudaMemcpyAsync(t, ctx->dTrans.p, size_t(ctx->nJobs) * 8, cudaMemcpyDeviceToHost, ctx->L.stream));
      CK(cudaStreamSynchronize(ctx->L.stream));
    }
    return MCI_OK;
  }

} // extern "C"
