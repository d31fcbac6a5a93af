// mci_types.h -- POD records shared by the host side of the library and its kernels.
// Vocabulary follows the reference: slice = 2-D cross-section perpendicular to an
// axis; segment = (axis, label, slice i, slice j); job = one Align + interpolation tree;
// node = one Interpolate1to1 call (reference X:556-793).
#pragma once
#include <stdint.h>

namespace mci
{

typedef unsigned long long u64;

// Words (32 pixels each) per histogram work item of the distance-transform median: one warp searches the nearest
// intersection pixel for every (i xor j) pixel of the item, and a level lasts as long as its slowest item, so items are
// kept small (2 words = at most two pixels per lane) -- the host sizes the item list with the same constant.
#define MCI_DT_HCHUNK 2

// A bit-packed 2-D mask living in the mask arena.  Pixel (x, y) of the slice, with
// x0 <= x < x0+w and y0 <= y < y0+h, is bit ((x-x0)&31) of word off + (y-y0)*pitch + ((x-x0)>>5).
// Bits at or beyond w in the last word of a row are always 0.
struct MaskRef
{
  int x0, y0, w, h;
  int pitch; // 32-bit words per row
  int pad;
  u64 off; // word offset in the arena
};

// One annotated slice of the batch handed to mci_label_slices.
struct SliceDev
{
  int axis, slice, label;
  int u0, v0, w, h;
  int runCap;    // capacity of this slice's run table
  u64 pixOff;    // offset into the component-image array
  u64 bitOff;    // offset (words) into the bit-row / per-word run-number arrays
  u64 runOff;    // offset into the run -> component table
  u64 base;      // voxel index of slice pixel (u0, v0)
  long long su, sv; // voxel strides of u and v
};

// A tile of rows of one slice (CC kernels) or one segment (pair kernel).
struct RowTile
{
  int item; // slice index or segment index
  int row0, nrows;
};

struct SegDev
{
  int si, sj; // slice indices
};

struct CompStat
{
  long long count;
  long long sumU, sumV;
  int minU, minV, maxU, maxV;
};

struct ExtractJob
{
  int slice, id;
  int row0, nrows; // tile of the mask's rows
  MaskRef m;
};

enum
{
  JOB_ONE_TO_ONE = 0,
  JOB_ONE_TO_N = 1,
  JOB_EXTRAPOLATE = 2
};

struct AlignJob
{
  int kind;
  int nB;
  int aRef; // index into the MaskRef table
  int bRef; // first of nB consecutive MaskRefs
  int indx, indy; // centroid offset (second BFS seed)
  int sx0, sy0, sw, sh; // search region of translations
  int phx, phy; // phantom pixel (EXTRAPOLATE)
  int qcap; // ring capacity (power of two)
  int useSmem;
  int out;  // index of the job this translation belongs to (jobs are launched in two groups)
  int pad0;
  long long minIter, maxIter;
  u64 bitmapOff; // global scratch (words) when !useSmem
  u64 queueOff;  // global scratch (words) when !useSmem: qcap*3 words
};

struct SplitJob
{
  int job;     // index of the AlignJob whose translation is used
  int aRef;    // mask being split
  int nB, bRef;
  int outRef;  // first of nB output MaskRefs (geometry of A)
  int pad;
  u64 scratchOff; // nB planes of A.h*A.pitch words (ping-pong buffer)
};

// One Interpolate1to1 call: shapes A (slice position i) and B (slice position j), with
// translation t such that pixel p of A corresponds to pixel p+t of B (reference X:556-568).
struct Node
{
  MaskRef A, B;
  int tx, ty;
  int i, j;
  int axis, label;
  // trees whose slices are perpendicular to x (axis 0) register their mid slices in a slot table, one slot per slice
  // position between the root's two slices, so that the volume writes of a whole tree can go x-run by x-run
  // (k_apply_axis0): slot = slot0 + (mid - lo - 1); slot0 < 0: not registered
  int slot0, lo;
};

// Per-node state of one level of the distance-transform pipeline (mci_dt.cu).
struct NodeWork
{
  Node nd;
  int ux0, uy0, w, h, pitch, words; // working region = union of the two translated shapes' regions
  int iTx, iTy, jTx, jTy, mid;
  int xr0, xr1;                     // first / last row holding intersection pixels
  int bigIdx;                       // slot of the wrapped-bin table (-1: none)
  int thr;                          // d2 threshold of the median (-1: the median is the intersection)
  int bx0, by0, bx1, by1;           // bounding box of the clipped median
  int alive;
  u64 scratchOff;                   // Xb | Ib | Sb (words each) | rowInfo (h)
  MaskRef M;                        // the mid shape
};

// One mid slice produced by a run: enough to replay its volume write elsewhere (multi-GPU combine).
struct EmitRec
{
  MaskRef M; // off is relative to the first mid-shape word of the arena
  int axis, label, mid, pad;
};
// The axis-0 slot table (ints, record index or -1) lives behind the record list: emit + emitCap.
__host__ __device__ inline int *
emit_slot_table(EmitRec * emit, int emitCap)
{
  return reinterpret_cast<int *>(emit + emitCap);
}

struct DtItem
{
  int node;
  int w0; // first word of the chunk
};

struct RootJob
{
  int kind;
  int job;     // AlignJob index
  int aRef, bRef, nB;
  int splitRef; // ONE_TO_N: first of nB split masks
  int posA, posB;
  int axis, label;
  int rootOff; // index of the first root node written
  int phRef;   // EXTRAPOLATE: MaskRef slot (1 word) for the phantom
  int phx, phy;
  int slot0;   // first slot of the first root's tree in the axis-0 slot table (-1: none); further roots follow at gap-1 each
};

// up to 32 consecutive slots of one axis-0 tree (k_apply_axis0)
struct Axis0Group
{
  int slot0, n;
};

struct VolDev
{
  long long X, Y, Z;
};

struct DevParams
{
  int heuristic, useDT, useBall;
};

enum
{
  ERR_NONE = 0,
  ERR_ARENA = 1,
  ERR_NODE_LIST = 2,
  ERR_PAIR_TABLE = 3,
  ERR_COMPONENTS = 4,
  ERR_PIXELTYPE = 5,
  ERR_QUEUE = 6,
  ERR_SCRATCH = 7,
  ERR_SPLIT_STUCK = 8
};

} // namespace mci
