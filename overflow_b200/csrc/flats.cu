// flats.cu -- flat resolution (Barnes, Lehman & Mulla 2014) on the device: the reference's
// src/overflow/fix_flats.py (flat_edges :13-62, label_flats :65-108, away_from_higher :111-161,
// towards_lower :164-224, resolve_flats :227-288, d8_masked_flow_dirs :291-339).
//
// The reference is a chain of sequential scans and FIFO sweeps.  What it computes is order-free, and that is
// what runs here:
//   * flat_edges is a 3x3 stencil (a cell is a low edge iff it has a direction and an equally high
//     neighbour without one; a high edge iff it has none and a higher neighbour that is not NODATA).
//   * label_flats floods cells of equal elevation, 8-connected, whatever their direction code; labels are
//     handed out in the order the row-major low-edge list meets unlabelled cells.  Here: union-find over
//     "equal elevation" links (atomicMin hooking, path halving), the smallest low-edge index of every
//     component by atomicMin, and label = 1 + rank of that index among all components' smallest low edges
//     (a prefix sum over the raster).  Components without a low edge keep label 0.
//   * away_from_higher / towards_lower are breadth-first sweeps with a level marker: a cell's value is its
//     BFS level from the seed edges through same-label cells without a direction.  Here: level-synchronous
//     frontier queues (ballot-compacted appends, one claim per cell by compare-and-swap), the level loop
//     batched on the host with a count read back per batch; flat_height[label] = max level by atomicMax
//     (the reference's "last write" is the largest level because levels only grow).
//   * d8_masked_flow_dirs is a 3x3 stencil over flat_mask and labels, slopes in float64 as in the reference.
// Everything is integer work except the two float compares (==, <) on elevations and the float64 slope
// division, so results are compared bit for bit with the reference's.
#include <limits.h>
#include <math_constants.h>

#include "common.cuh"

namespace ofl {

namespace {

constexpr int FL_UNDEF = OFL_DIR_UNDEFINED, FL_NODATA = OFL_DIR_NODATA;
constexpr int FL_BIG = INT_MAX;
constexpr int FL_THREADS = 256;
constexpr int FL_SCAN_THREADS = 1024;

// counters (uint32) kept in device memory
enum { CNT_LOW = 0, CNT_HIGH, CNT_LABELS, CNT_SEEDS, CNT_FRONT0, CNT_FRONT1, CNT_FRONT2, CNT_SLOTS = 16 };

__constant__ int c_dy[8] = {0, -1, -1, -1, 0, 1, 1, 1};  // E NE N NW W SW S SE: constants.py:29-40
__constant__ int c_dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};

// ---------------------------------------------------------------- flat_edges (fix_flats.py:13-62)
// Also initialises the union-find forest and the per-component "smallest low edge" slot.  The forest starts with
// every horizontal run of equal elevation already hanging off its first cell as far as one warp sees it (a ballot
// and a count-leading-zeros, no atomics), which is most of the linking on a plateau.
__global__ void __launch_bounds__(FL_THREADS)
flat_edges_kernel(const float* __restrict__ dem, const uint8_t* __restrict__ fdr, int rows, int cols, uint8_t* edges,
                  int* parent, int* minlow, unsigned* cnt) {
  const unsigned n = (unsigned)rows * (unsigned)cols;
  const unsigned i = blockIdx.x * FL_THREADS + threadIdx.x;
  int flag = 0;
  bool left_eq = false;
  if (i < n) {
    const int r = (int)(i / (unsigned)cols), c = (int)(i - (unsigned)r * (unsigned)cols);
    const int cur = fdr[i];
    const float z = dem[i];
    left_eq = c > 0 && dem[i - 1] == z;
    // The two kinds exclude each other, so each cell runs one loop only, and the result does not depend on the
    // order the neighbours are visited in.  A cell with a direction reads a neighbour's elevation only where
    // that neighbour has none (rare outside flats); cells off the raster's ring skip the bounds checks.
    const bool inner = r > 0 && r + 1 < rows && c > 0 && c + 1 < cols;
    if (cur != FL_UNDEF) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int nr = r + c_dy[k], nc = c + c_dx[k];
        if (!inner && (nr < 0 || nr >= rows || nc < 0 || nc >= cols)) continue;
        const unsigned j = (unsigned)nr * (unsigned)cols + (unsigned)nc;
        if (fdr[j] == FL_UNDEF && dem[j] == z) flag = 1;  // :45-53
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int nr = r + c_dy[k], nc = c + c_dx[k];
        if (!inner && (nr < 0 || nr >= rows || nc < 0 || nc >= cols)) continue;
        const unsigned j = (unsigned)nr * (unsigned)cols + (unsigned)nc;
        if (z < dem[j] && fdr[j] != FL_NODATA) flag = 2;  // :41-43, :54-60
      }
    }
    edges[i] = (uint8_t)flag;
  }
  if (parent) {
    const int lane = threadIdx.x & 31;
    const unsigned em = __ballot_sync(0xffffffffu, left_eq);
    const unsigned starts = (~em & ((2u << lane) - 1u)) | 1u;  // lanes <= mine that begin a run (lane 0 always does)
    if (i < n) {
      parent[i] = (int)(i - (unsigned)(lane - (31 - __clz(starts))));
      minlow[i] = FL_BIG;
    }
  }
  const int lo = __syncthreads_count(flag == 1), hi = __syncthreads_count(flag == 2);
  if (threadIdx.x == 0) {
    if (lo) atomicAdd(&cnt[CNT_LOW], (unsigned)lo);
    if (hi) atomicAdd(&cnt[CNT_HIGH], (unsigned)hi);
  }
}

// ---------------------------------------------------------------- equal-elevation components
__device__ __forceinline__ int uf_find(int* p, int i) {
  int cur = __ldcg(p + i);
  while (cur != i) {
    const int nxt = __ldcg(p + cur);
    if (nxt != cur) p[i] = nxt;  // path halving; any ancestor is a valid parent
    i = cur;
    cur = nxt;
  }
  return i;
}

// read-only walk to the root (used once the forest is final: concurrent halving could overwrite a flattened entry)
__device__ __forceinline__ int uf_find_ro(const int* p, int i) {
  int cur = __ldcg(p + i);
  while (cur != i) {
    i = cur;
    cur = __ldcg(p + cur);
  }
  return i;
}

__device__ __forceinline__ void uf_unite(int* p, int a, int b) {
  for (;;) {
    a = uf_find(p, a);
    b = uf_find(p, b);
    if (a == b) return;
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    const int old = atomicMin(p + a, b);  // hook the larger root under the smaller one
    if (old == a) return;
    a = old;  // a was no longer a root: carry on from what it pointed to
  }
}

// Unions between runs.  Inside a row only the hand-over between warps is left (lane 31 -> lane 0).  Between rows a
// vertical pair needs a union only where one of the two runs begins (further right the pair to the left is the same
// two runs), and a diagonal pair only when neither the cell below nor the cell beside it continues a run that the
// vertical rule already ties in.
__global__ void __launch_bounds__(FL_THREADS)
flat_merge_kernel(const float* __restrict__ dem, int rows, int cols, int* parent) {
  const unsigned n = (unsigned)rows * (unsigned)cols;
  const unsigned i = blockIdx.x * FL_THREADS + threadIdx.x;
  if (i >= n) return;
  const int r = (int)(i / (unsigned)cols), c = (int)(i - (unsigned)r * (unsigned)cols);
  const float z = dem[i];
  const bool left_eq = c > 0 && dem[i - 1] == z;
  const bool right_eq = c + 1 < cols && dem[i + 1] == z;
  if (right_eq && (threadIdx.x & 31) == 31) uf_unite(parent, (int)i, (int)i + 1);
  if (r + 1 < rows) {
    const unsigned j = i + (unsigned)cols;
    if (dem[j] == z) {
      if (!left_eq || !(dem[j - 1] == z)) uf_unite(parent, (int)i, (int)j);  // c == 0 implies !left_eq
    } else {
      if (c > 0 && !left_eq && dem[j - 1] == z) uf_unite(parent, (int)i, (int)(j - 1));
      if (c + 1 < cols && !right_eq && dem[j + 1] == z) uf_unite(parent, (int)i, (int)(j + 1));
    }
  }
}

// The passes below only care about the flagged few cells: a thread reads four flag bytes as one word and a
// 1024-thread CTA covers 4096 cells (one thread per byte spent its time scheduling CTAs, not moving data).
constexpr int FL_CELLS_PER_CTA = 4 * FL_SCAN_THREADS;

__device__ __forceinline__ unsigned flag_word(const uint8_t* bytes, int64_t n, int64_t i0, bool aligned) {
  if (i0 >= n) return 0u;
  if (aligned && i0 + 4 <= n) return *reinterpret_cast<const unsigned*>(bytes + i0);
  unsigned w = 0;
  for (int k = 0; k < 4; ++k)
    if (i0 + k < n) w |= (unsigned)bytes[i0 + k] << (8 * k);
  return w;
}

// exclusive prefix of a per-thread count over the CTA (1024 threads); returns the CTA total through *total
__device__ __forceinline__ unsigned cta_exclusive(unsigned c, unsigned* total) {
  __shared__ unsigned warp_sum[FL_SCAN_THREADS / 32];
  __shared__ unsigned cta_total;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned inc = c;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sum[w] = inc;
  __syncthreads();
  if (w == 0) {
    const unsigned v = warp_sum[lane];
    unsigned ws = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, ws, off);
      if (lane >= off) ws += t;
    }
    warp_sum[lane] = ws - v;
    if (lane == 31) cta_total = ws;
  }
  __syncthreads();
  *total = cta_total;
  return warp_sum[w] + inc - c;
}

// low edges look their root up (making their own pointer direct) and post their index to the root's slot
__global__ void __launch_bounds__(FL_SCAN_THREADS)
flat_lowroot_kernel(int64_t n, int* parent, const uint8_t* __restrict__ edges, int* minlow) {
  const int64_t i0 = 4 * ((int64_t)blockIdx.x * FL_SCAN_THREADS + threadIdx.x);
  const unsigned w = flag_word(edges, n, i0, true) & 0x01010101u;
  if (w == 0) return;
  for (int k = 0; k < 4; ++k)
    if ((w >> (8 * k)) & 1u) {
      const int i = (int)(i0 + k);
      const int root = uf_find_ro(parent, i);
      parent[i] = root;  // own entry only, and a root is a valid parent for any reader walking by
      if (__ldcg(minlow + root) > i) atomicMin(minlow + root, i);  // a plain look first: most low edges lose to an earlier one
    }
}

// a seed is the first low edge (row-major) of its component; bit k of the result: cell i0 + k is one
__device__ __forceinline__ unsigned seed_bits(int64_t n, int64_t i0, const int* parent, const uint8_t* edges,
                                              const int* minlow) {
  const unsigned w = flag_word(edges, n, i0, true) & 0x01010101u;
  unsigned bits = 0;
  if (w)
    for (int k = 0; k < 4; ++k)
      if (((w >> (8 * k)) & 1u) && minlow[parent[i0 + k]] == (int)(i0 + k)) bits |= 1u << k;
  return bits;
}

__global__ void __launch_bounds__(FL_SCAN_THREADS)
flat_seed_count_kernel(int64_t n, const int* parent, const uint8_t* __restrict__ edges, const int* minlow, int* blockcnt) {
  const int64_t i0 = 4 * ((int64_t)blockIdx.x * FL_SCAN_THREADS + threadIdx.x);
  unsigned total;
  cta_exclusive(__popc(seed_bits(n, i0, parent, edges, minlow)), &total);
  if (threadIdx.x == 0) blockcnt[blockIdx.x] = (int)total;
}

// exclusive scan of blockcnt[nb] in place (one CTA), total -> cnt[CNT_LABELS]
__global__ void __launch_bounds__(FL_SCAN_THREADS) flat_seed_scan_kernel(int nb, int* blockcnt, unsigned* cnt) {
  __shared__ int part[FL_SCAN_THREADS];
  const int t = threadIdx.x;
  const int per = (nb + FL_SCAN_THREADS - 1) / FL_SCAN_THREADS;
  const int lo = t * per, hi = min(nb, lo + per);
  int s = 0;
  for (int k = lo; k < hi; ++k) s += blockcnt[k];
  part[t] = s;
  __syncthreads();
  for (int off = 1; off < FL_SCAN_THREADS; off <<= 1) {  // Hillis-Steele inclusive scan
    const int v = t >= off ? part[t - off] : 0;
    __syncthreads();
    part[t] += v;
    __syncthreads();
  }
  int run = part[t] - s;
  for (int k = lo; k < hi; ++k) {
    const int v = blockcnt[k];
    blockcnt[k] = run;
    run += v;
  }
  if (t == FL_SCAN_THREADS - 1) cnt[CNT_LABELS] = (unsigned)part[t];
}

// seedlabel[i] = 1 + rank of seed i (row-major)
__global__ void __launch_bounds__(FL_SCAN_THREADS)
flat_seed_rank_kernel(int64_t n, const int* parent, const uint8_t* __restrict__ edges, const int* minlow,
                      const int* __restrict__ blockoff, int* seedlabel) {
  const int64_t i0 = 4 * ((int64_t)blockIdx.x * FL_SCAN_THREADS + threadIdx.x);
  const unsigned bits = seed_bits(n, i0, parent, edges, minlow);
  unsigned total;
  unsigned before = cta_exclusive(__popc(bits), &total) + (unsigned)blockoff[blockIdx.x];
  for (int k = 0; k < 4; ++k)
    if ((bits >> k) & 1u) seedlabel[i0 + k] = (int)(++before);
}

// Every cell takes the label of its component: root -> the component's smallest low edge -> that seed's label
// (0 when the component has no low edge).  The walk to the root is read-only (low edges already point at it).
// open[i] is the sweeps' one-word view of a cell: 0 = never a candidate (it has a direction), otherwise
// (label + 1) << 2 | g with g = 0 untouched, 1 claimed by the away sweep, 2 claimed by the towards sweep.
__device__ __forceinline__ unsigned flat_state(int lab) { return (unsigned)(lab + 1) << 2; }

__global__ void __launch_bounds__(FL_THREADS)
flat_spread_label_kernel(int64_t n, const int* __restrict__ parent, const int* __restrict__ minlow,
                         const int* __restrict__ seedlabel, int* labels, const uint8_t* __restrict__ fdr, unsigned* open) {
  const int64_t i = (int64_t)blockIdx.x * FL_THREADS + threadIdx.x;
  if (i >= n) return;
  const int m = minlow[uf_find_ro(parent, (int)i)];
  const int lab = (m == FL_BIG) ? 0 : seedlabel[m];
  labels[i] = lab;
  open[i] = (fdr[i] == FL_UNDEF) ? flat_state(lab) : 0u;
}

// ---------------------------------------------------------------- gradients (away_from_higher / towards_lower)
// edge cells -> seed list, in no particular order (the sweeps are order-free); one queue atomic per CTA
__global__ void __launch_bounds__(FL_SCAN_THREADS)
flat_collect_kernel(int64_t n, const uint8_t* __restrict__ edges, unsigned bit, const int* __restrict__ labels,
                    int* seeds, unsigned* cnt) {
  __shared__ unsigned cta_base;
  const int64_t i0 = 4 * ((int64_t)blockIdx.x * FL_SCAN_THREADS + threadIdx.x);
  const unsigned w = flag_word(edges, n, i0, true) & (bit * 0x01010101u);
  unsigned bits = 0;
  if (w)
    for (int k = 0; k < 4; ++k)  // fix_flats.py:273-274: high edges outside labelled flats drop out (low edges never do)
      if ((w >> (8 * k)) & 0xffu && labels[i0 + k] != 0) bits |= 1u << k;
  unsigned total;
  unsigned at = cta_exclusive(__popc(bits), &total);
  if (total == 0) return;
  if (threadIdx.x == 0) cta_base = atomicAdd(&cnt[CNT_SEEDS], total);
  __syncthreads();
  at += cta_base;
  for (int k = 0; k < 4; ++k)
    if ((bits >> k) & 1u) seeds[at++] = (int)(i0 + k);
}

// Sweep modes: 0 away_from_higher; 1 towards_lower on a mask that was negated first (the standalone entry point,
// any caller-supplied mask); 2 towards_lower straight on the away sweep's mask (resolve_flats: every cell the away
// sweep valued is reached again, so the reference's negation never shows in the result and is not materialised).
enum { SWEEP_AWAY = 0, SWEEP_TOWARDS_NEGATED = 1, SWEEP_TOWARDS = 2 };

// standalone sweeps: towards_lower negates the mask (fix_flats.py:200); candidates are armed from the mask
__global__ void __launch_bounds__(FL_THREADS)
flat_prepare_kernel(int64_t n, int mode, int* flat_mask, const int* __restrict__ labels, const uint8_t* __restrict__ fdr,
                    unsigned* open) {
  const int64_t i = (int64_t)blockIdx.x * FL_THREADS + threadIdx.x;
  if (i >= n) return;
  int fm = flat_mask[i];
  if (mode != SWEEP_AWAY) {
    fm = -fm;
    flat_mask[i] = fm;
  }
  open[i] = (fdr[i] == FL_UNDEF && fm <= 0) ? flat_state(labels[i]) : 0u;
}

// One try to claim candidate word `s` of cell q for this sweep.
__device__ __forceinline__ bool flat_claim(unsigned* open, int q, unsigned s, unsigned want, int mode) {
  if ((s & ~3u) != want) return false;
  if (mode == SWEEP_AWAY) return (s & 3u) == 0u && atomicCAS(open + q, s, s | 1u) == s;
  return (s & 3u) != 2u && atomicCAS(open + q, s, (s & ~3u) | 2u) == s;
}

// The value a cell gets when a sweep reaches it at `level` (:153-154 / :209-214); only the claiming thread calls this.
__device__ __forceinline__ void flat_assign(int p, int level, int mode, int lab, int* flat_mask, const int* fh_read,
                                            int* fh_acc) {
  if (mode == SWEEP_AWAY) {
    flat_mask[p] = level;
    // every cell of a level carries the same value: one atomic per flat and level gets through
    if (lab > 0 && __ldcg(fh_acc + lab - 1) < level) atomicMax(fh_acc + lab - 1, level);
  } else {
    const int fm = flat_mask[p];
    int away = 0;  // flat_height - (increments away from higher terrain), 0 where the away sweep never came
    if (lab > 0) {
      if (mode == SWEEP_TOWARDS_NEGATED && fm < 0) away = fm + fh_read[lab - 1];
      if (mode == SWEEP_TOWARDS && fm > 0) away = fh_read[lab - 1] - fm;
    }
    flat_mask[p] = away + 2 * level;
  }
}

// Frontier appends go through a per-CTA buffer: a warp reserves its slots with one shared-memory atomic, and the
// CTA moves the buffer to the global queue with one global atomic when it fills up.  (One global atomic per warp
// was the bottleneck of the sweeps: tens of millions of adds on a single counter.)
#ifndef OFL_FL_ILP
#define OFL_FL_ILP 2
#endif
constexpr int FL_ILP = OFL_FL_ILP;  // pushes per thread and loop round (tuning: -DOFL_FL_ILP=4)
constexpr int FL_QBUF = 1024 * FL_ILP;
constexpr int FL_ROUNDS = 2;  // loop rounds between two looks at the fill level
constexpr int FL_BATCH = FL_ROUNDS * FL_ILP * FL_THREADS;  // most pushes between two looks
static_assert(FL_QBUF >= 2 * FL_BATCH, "buffer must hold two batches of rounds");

struct BlockQueue {
  int buf[FL_QBUF];
  unsigned n, base;
};

__device__ __forceinline__ void bq_push(BlockQueue& bq, bool won, int p) {
  const unsigned m = __ballot_sync(0xffffffffu, won);
  if (m == 0) return;
  const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
  unsigned off = 0;
  if (lane == leader) off = atomicAdd(&bq.n, (unsigned)__popc(m));
  off = __shfl_sync(0xffffffffu, off, leader);
  if (won) bq.buf[off + __popc(m & ((1u << lane) - 1u))] = p;
}

// all threads of the CTA; pushes must not overlap a flush
__device__ __forceinline__ void bq_flush(BlockQueue& bq, int* qout, unsigned* cnt_out) {
  __syncthreads();
  const unsigned n = bq.n;
  if (n == 0) return;  // uniform: read between two barriers with no push in flight
  if (threadIdx.x == 0) bq.base = atomicAdd(cnt_out, n);
  __syncthreads();
  const unsigned base = bq.base;
  for (unsigned k = threadIdx.x; k < n; k += FL_THREADS) qout[base + k] = bq.buf[k];
  __syncthreads();
  if (threadIdx.x == 0) bq.n = 0;
  __syncthreads();
}

// after every FL_ROUNDS rounds: flush when another batch of rounds might not fit
__device__ __forceinline__ void bq_maybe_flush(BlockQueue& bq, unsigned round, int* qout, unsigned* cnt_out) {
  if (round % FL_ROUNDS) return;
  __syncthreads();
  const bool full = bq.n > FL_QBUF - FL_BATCH;
  __syncthreads();
  if (full) bq_flush(bq, qout, cnt_out);
}

// level 1: the seed edges themselves (duplicates and already positive cells drop out)
__global__ void __launch_bounds__(FL_THREADS)
flat_seed_level_kernel(const int* __restrict__ seeds, const int* __restrict__ labels, int mode, int* flat_mask,
                       unsigned* open, const int* fh_read, int* fh_acc, int* qout, unsigned* cnt) {
  __shared__ BlockQueue bq;
  const unsigned n_seed = cnt[CNT_SEEDS];
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[CNT_FRONT0 + 2] = 0;
  if (threadIdx.x == 0) bq.n = 0;
  __syncthreads();
  const unsigned stride = gridDim.x * FL_THREADS;
  unsigned round = 0;
  for (unsigned base = blockIdx.x * FL_THREADS; base < n_seed; base += stride) {
    const unsigned idx = base + threadIdx.x;
    bool won = false;
    int p = 0;
    if (idx < n_seed) {
      p = seeds[idx];
      const int lab = labels[p];
      const unsigned o = __ldcg(open + p);
      if (o != 0u) {  // a candidate cell: claimed like any other
        won = flat_claim(open, p, o, flat_state(lab), mode);
        if (won) flat_assign(p, 1, mode, lab, flat_mask, fh_read, fh_acc);
      } else {  // a cell with a direction (low edges) or one that already has its value: the mask itself dedupes
        const int fm = __ldcg(flat_mask + p);
        if (fm <= 0) {
          const int nv = mode == SWEEP_AWAY ? 1 : ((fm < 0 && lab > 0) ? fm + fh_read[lab - 1] : 0) + 2;
          won = atomicCAS(flat_mask + p, fm, nv) == fm;
          if (won && mode == SWEEP_AWAY && lab > 0) atomicMax(fh_acc + lab - 1, 1);
        }
      }
    }
    bq_push(bq, won, p);
    bq_maybe_flush(bq, ++round, qout, &cnt[CNT_FRONT0 + 1]);
  }
  bq_flush(bq, qout, &cnt[CNT_FRONT0 + 1]);
}

// level L -> L+1: neighbours with the same label and no direction (:155-161 / :215-224).  Eight lanes share a
// frontier cell, one neighbour each: a level is one load-compare-claim deep instead of eight (small frontiers are
// latency bound), and a neighbour costs one scattered word (open[q]) instead of three (mask, code, label).
__global__ void __launch_bounds__(FL_THREADS)
flat_level_kernel(int level, const int* __restrict__ qin, int* qout, const int* __restrict__ labels, int rows, int cols,
                  int mode, int* flat_mask, unsigned* open, const int* fh_read, int* fh_acc, unsigned* cnt) {
  __shared__ BlockQueue bq;
  const unsigned n_in = cnt[CNT_FRONT0 + level % 3];
  unsigned* cnt_out = &cnt[CNT_FRONT0 + (level + 1) % 3];
  if (blockIdx.x == 0 && threadIdx.x == 0) cnt[CNT_FRONT0 + (level + 2) % 3] = 0;
  if (threadIdx.x == 0) bq.n = 0;
  __syncthreads();
  constexpr unsigned CELLS = FL_THREADS / 8;
  const unsigned stride = gridDim.x * CELLS * FL_ILP;
  const int k = threadIdx.x & 7;
  const int dy = c_dy[k], dx = c_dx[k];
  unsigned round = 0;
  for (unsigned base = blockIdx.x * CELLS * FL_ILP; base < n_in; base += stride) {
    // FL_ILP independent (cell, neighbour) pairs per thread: the three dependent loads of each overlap
    int q[FL_ILP];
    unsigned want[FL_ILP], seen[FL_ILP];
    bool ok[FL_ILP];
#pragma unroll
    for (int u = 0; u < FL_ILP; ++u) {
      const unsigned idx = base + u * CELLS + (threadIdx.x >> 3);
      ok[u] = idx < n_in;
      q[u] = ok[u] ? qin[idx] : 0;
    }
#pragma unroll
    for (int u = 0; u < FL_ILP; ++u) {
      const int p = q[u];
      want[u] = ok[u] ? flat_state(labels[p]) : 0u;
      const int r = p / cols, c = p - r * cols;
      const int nr = r + dy, nc = c + dx;
      ok[u] = ok[u] && nr >= 0 && nr < rows && nc >= 0 && nc < cols;
      q[u] = ok[u] ? nr * cols + nc : 0;
    }
#pragma unroll
    for (int u = 0; u < FL_ILP; ++u) seen[u] = ok[u] ? __ldcg(open + q[u]) : 0u;
#pragma unroll
    for (int u = 0; u < FL_ILP; ++u) {
      bool won = false;
      if (ok[u] && seen[u] != 0u && flat_claim(open, q[u], seen[u], want[u], mode)) {
        won = true;
        flat_assign(q[u], level + 1, mode, (int)(want[u] >> 2) - 1, flat_mask, fh_read, fh_acc);
      }
      bq_push(bq, won, q[u]);
    }
    bq_maybe_flush(bq, ++round, qout, cnt_out);
  }
  bq_flush(bq, qout, cnt_out);
}

// flat_height[k] takes the sweep's value where the sweep reached label k+1 (standalone away_from_higher)
__global__ void __launch_bounds__(FL_THREADS) flat_height_merge_kernel(int n, const int* __restrict__ acc, int* fh) {
  const int i = blockIdx.x * FL_THREADS + threadIdx.x;
  if (i < n && acc[i] > 0) fh[i] = acc[i];
}

// ---------------------------------------------------------------- d8_masked_flow_dirs (fix_flats.py:291-339)
// Slopes in float64 with the reference's division (:335-340).  (An exact integer comparison, b^2 against 2 a^2,
// was measured: it needs all eight differences in registers at once and ran 18 % slower; the kernel waits on its
// sixteen neighbour loads, not on the divisions.)
__device__ __forceinline__ int masked_dir_of(int64_t i, const int* __restrict__ flat_mask, const int* __restrict__ labels,
                                             int rows, int cols) {
  const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
  const int lab = labels[i];
  const double fm = (double)flat_mask[i];
  int nmin = FL_UNDEF;
  double min_slope = CUDART_INF;
  const bool inner = r > 0 && r + 1 < rows && c > 0 && c + 1 < cols;  // off the raster's ring: no bounds checks
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int nr = r + c_dy[k], nc = c + c_dx[k];
    if (!inner && (nr < 0 || nr >= rows || nc < 0 || nc >= cols)) continue;
    const int64_t j = (int64_t)nr * cols + nc;
    if (labels[j] != lab) continue;
    const double dz = (double)flat_mask[j] - fm;
    const double slope = (k & 1) ? __ddiv_rn(dz, 1.4142135623730951) : dz;  // odd codes are the diagonals
    if (slope < min_slope) {
      min_slope = slope;
      nmin = k;
    }
  }
  return nmin;
}

// a thread looks at four codes at once; only cells without a direction cost anything
__global__ void __launch_bounds__(FL_THREADS)
flat_masked_dirs_kernel(const int* __restrict__ flat_mask, const int* __restrict__ labels, uint8_t* fdr, int rows,
                        int cols, int aligned) {
  const int64_t n = (int64_t)rows * cols;
  const int64_t i0 = 4 * ((int64_t)blockIdx.x * FL_THREADS + threadIdx.x);
  const unsigned w = flag_word(fdr, n, i0, aligned != 0);
  for (int k = 0; k < 4; ++k)
    if (i0 + k < n && ((w >> (8 * k)) & 0xffu) == (unsigned)FL_UNDEF)
      fdr[i0 + k] = (uint8_t)masked_dir_of(i0 + k, flat_mask, labels, rows, cols);
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

struct FlatsWork {
  int* parent;  // union-find forest, later flat_height
  int* q0;      // frontier queue / block counts of the label scan
  int* q1;      // seed list / frontier queue
  unsigned* open;  // the sweeps' candidate words
  uint8_t* edges;
  unsigned* cnt;
};

size_t flats_align(size_t v) { return (v + 255) / 256 * 256; }

int carve(void* workspace, size_t workspace_bytes, int64_t n, FlatsWork* w) {
  const size_t a = flats_align((size_t)n * sizeof(int)), e = flats_align((size_t)n), c = flats_align(CNT_SLOTS * sizeof(unsigned));
  OFL_REQUIRE(workspace_bytes >= 4 * a + e + c, OFL_ERR_WORKSPACE, "flats workspace too small: %zu < %zu", workspace_bytes,
              4 * a + e + c);
  char* p = static_cast<char*>(workspace);
  w->parent = reinterpret_cast<int*>(p);
  w->q0 = reinterpret_cast<int*>(p + a);
  w->q1 = reinterpret_cast<int*>(p + 2 * a);
  w->open = reinterpret_cast<unsigned*>(p + 3 * a);
  w->edges = reinterpret_cast<uint8_t*>(p + 4 * a);
  w->cnt = reinterpret_cast<unsigned*>(p + 4 * a + e);
  return OFL_OK;
}

// One sweep from the seed list in w.q1 (count in cnt[CNT_SEEDS]).  Host-synchronous: the level loop reads the
// next frontier's size back once per batch of launches.
int run_gradient(int rows, int cols, const int* labels, const uint8_t* fdr, int mode, bool open_ready, int* flat_mask,
                 const int* fh_read, int* fh_acc, const FlatsWork& w, int64_t* levels_out, cudaStream_t st) {
  // open_ready: w.open already describes the candidates (resolve_flats: armed by the labelling, re-used by the
  // second sweep through the generation bits); otherwise they are armed from the mask, negated first for towards_lower
  const int64_t n = (int64_t)rows * cols;
  OFL_CUDA(cudaMemsetAsync(w.cnt + CNT_FRONT0, 0, 3 * sizeof(unsigned), st));
  if (!open_ready) {
    flat_prepare_kernel<<<blocks_for(n, FL_THREADS), FL_THREADS, 0, st>>>(n, mode, flat_mask, labels, fdr, w.open);
    OFL_CHECK_LAUNCH();
  }
  const unsigned grid = (unsigned)(sm_count() * 8);  // 8 CTAs of 256 threads fill an SM
  flat_seed_level_kernel<<<grid, FL_THREADS, 0, st>>>(w.q1, labels, mode, flat_mask, w.open, fh_read, fh_acc, w.q0, w.cnt);
  OFL_CHECK_LAUNCH();
  unsigned* h_cnt = nullptr;
  int rc = pinned_get(64, reinterpret_cast<void**>(&h_cnt));
  if (rc != OFL_OK) return rc;
  int level = 1;  // the frontier in `qin` holds the cells of this level
  int* qin = w.q0;
  int* qout = w.q1;
  int batch = 8;
  for (;;) {
    for (int b = 0; b < batch; ++b) {
      flat_level_kernel<<<grid, FL_THREADS, 0, st>>>(level, qin, qout, labels, rows, cols, mode, flat_mask, w.open, fh_read,
                                                      fh_acc, w.cnt);
      OFL_CHECK_LAUNCH();
      ++level;
      int* t = qin;
      qin = qout;
      qout = t;
    }
    OFL_CUDA(cudaMemcpyAsync(h_cnt, w.cnt + CNT_FRONT0 + level % 3, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    OFL_CUDA(cudaStreamSynchronize(st));
    if (*h_cnt == 0) break;
    OFL_REQUIRE(level < INT_MAX / 4, OFL_ERR_INVALID, "flat gradient did not terminate");
    if (batch < 64) batch *= 2;
  }
  if (levels_out) *levels_out = level;
  return OFL_OK;
}

}  // namespace

size_t flats_workspace_bytes(int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  return 4 * flats_align((size_t)n * sizeof(int)) + flats_align((size_t)n) + flats_align(CNT_SLOTS * sizeof(unsigned));
}

static int check_shape(int64_t rows, int64_t cols) {
  OFL_REQUIRE(rows > 0 && cols > 0 && rows * cols < (int64_t)INT_MAX, OFL_ERR_INVALID,
              "flat resolution works on one tile of fewer than 2^31 cells (got %lld x %lld)", (long long)rows,
              (long long)cols);
  return OFL_OK;
}

// edges[i]: bit 0 low edge, bit 1 high edge; counts -> host.  Dense rasters (ld == cols), device pointers.
int launch_flat_edges(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, uint8_t* edges, int64_t* n_low,
                      int64_t* n_high, unsigned* cnt_dev, cudaStream_t st) {
  int rc = check_shape(rows, cols);
  if (rc != OFL_OK) return rc;
  const int64_t n = rows * cols;
  PhaseScope ps(PHASE_FLATS, st);
  OFL_CUDA(cudaMemsetAsync(cnt_dev, 0, CNT_SLOTS * sizeof(unsigned), st));
  flat_edges_kernel<<<blocks_for(n, FL_THREADS), FL_THREADS, 0, st>>>(dem, fdr, (int)rows, (int)cols, edges, nullptr,
                                                                      nullptr, cnt_dev);
  OFL_CHECK_LAUNCH();
  unsigned* h = nullptr;
  rc = pinned_get(64, reinterpret_cast<void**>(&h));
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(h, cnt_dev, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  if (n_low) *n_low = h[CNT_LOW];
  if (n_high) *n_high = h[CNT_HIGH];
  return OFL_OK;
}

// resolve_flats (fix_flats.py:227-288).  All pointers device, dense.  info (host, nullable) receives
// {low edges, high edges, labels, levels of the away sweep, levels of the towards sweep}.
int launch_resolve_flats(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, int* flat_mask, int* labels,
                         int64_t* info, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  int rc = check_shape(rows, cols);
  if (rc != OFL_OK) return rc;
  const int64_t n = rows * cols;
  FlatsWork w;
  rc = carve(workspace, workspace_bytes, n, &w);
  if (rc != OFL_OK) return rc;
  const unsigned nb = blocks_for(n, FL_THREADS), nbc = blocks_for(n, FL_CELLS_PER_CTA);
  OFL_CUDA(cudaMemsetAsync(w.cnt, 0, CNT_SLOTS * sizeof(unsigned), st));
  int* flat_height = w.parent;  // lives where the forest was once the labels are known (label count <= cell count)
  int* minlow = w.q1;           // per-component smallest low edge, until the labels are known
  int* seedlabel = flat_mask;   // label of each component's first low edge, until the labels are known
  unsigned* h = nullptr;
  rc = pinned_get(64, reinterpret_cast<void**>(&h));
  if (rc != OFL_OK) return rc;
  int64_t lv_away = 0, lv_low = 0;
  {
    PhaseScope ps(PHASE_FLATS_LABEL, st);
    flat_edges_kernel<<<nb, FL_THREADS, 0, st>>>(dem, fdr, (int)rows, (int)cols, w.edges, w.parent, minlow, w.cnt);
    OFL_CHECK_LAUNCH();
    flat_merge_kernel<<<nb, FL_THREADS, 0, st>>>(dem, (int)rows, (int)cols, w.parent);
    OFL_CHECK_LAUNCH();
    flat_lowroot_kernel<<<nbc, FL_SCAN_THREADS, 0, st>>>(n, w.parent, w.edges, minlow);
    OFL_CHECK_LAUNCH();
    flat_seed_count_kernel<<<nbc, FL_SCAN_THREADS, 0, st>>>(n, w.parent, w.edges, minlow, w.q0);
    OFL_CHECK_LAUNCH();
    flat_seed_scan_kernel<<<1, FL_SCAN_THREADS, 0, st>>>((int)nbc, w.q0, w.cnt);
    OFL_CHECK_LAUNCH();
    flat_seed_rank_kernel<<<nbc, FL_SCAN_THREADS, 0, st>>>(n, w.parent, w.edges, minlow, w.q0, seedlabel);
    OFL_CHECK_LAUNCH();
    // one look at the counts: nothing to do without low edges (fix_flats.py:258-264), and flat_height needs
    // only as many cleared entries as there are labels
    OFL_CUDA(cudaMemcpyAsync(h, w.cnt, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    OFL_CUDA(cudaStreamSynchronize(st));
    const unsigned n_low = h[CNT_LOW], n_high = h[CNT_HIGH], n_labels = h[CNT_LABELS];
    if (info) {
      info[0] = n_low;
      info[1] = n_high;
      info[2] = n_labels;
    }
    OFL_REQUIRE(n_labels < (1u << 29), OFL_ERR_INVALID, "more than 2^29 flats in one tile");
    if (n_labels == 0) {
      OFL_CUDA(cudaMemsetAsync(labels, 0, (size_t)n * sizeof(int), st));
      OFL_CUDA(cudaMemsetAsync(flat_mask, 0, (size_t)n * sizeof(int), st));
      return OFL_OK;
    }
    flat_spread_label_kernel<<<nb, FL_THREADS, 0, st>>>(n, w.parent, minlow, seedlabel, labels, fdr, w.open);
    OFL_CHECK_LAUNCH();
    OFL_CUDA(cudaMemsetAsync(flat_mask, 0, (size_t)n * sizeof(int), st));
    OFL_CUDA(cudaMemsetAsync(flat_height, 0, ((size_t)n_labels + 1) * sizeof(int), st));
  }
  PhaseScope ps(PHASE_FLATS_SWEEP, st);
  flat_collect_kernel<<<nbc, FL_SCAN_THREADS, 0, st>>>(n, w.edges, 2u, labels, w.q1, w.cnt);
  OFL_CHECK_LAUNCH();
  rc = run_gradient((int)rows, (int)cols, labels, fdr, SWEEP_AWAY, true, flat_mask, flat_height, flat_height, w, &lv_away, st);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemsetAsync(w.cnt + CNT_SEEDS, 0, sizeof(unsigned), st));
  flat_collect_kernel<<<nbc, FL_SCAN_THREADS, 0, st>>>(n, w.edges, 1u, labels, w.q1, w.cnt);
  OFL_CHECK_LAUNCH();
  rc = run_gradient((int)rows, (int)cols, labels, fdr, SWEEP_TOWARDS, true, flat_mask, flat_height, flat_height, w, &lv_low, st);
  if (rc != OFL_OK) return rc;
  if (info) {
    info[3] = lv_away;
    info[4] = lv_low;
  }
  return OFL_OK;
}

// Standalone away_from_higher (towards == 0) / towards_lower (towards == 1) from a caller-supplied seed list
// (cell indices, device).  flat_height has n_heights entries (device).
int launch_flat_gradient(const int* labels, const uint8_t* fdr, int64_t rows, int64_t cols, const int* seeds,
                         int64_t n_seeds, int towards, int* flat_mask, int* flat_height, int64_t n_heights,
                         void* workspace, size_t workspace_bytes, cudaStream_t st) {
  int rc = check_shape(rows, cols);
  if (rc != OFL_OK) return rc;
  const int64_t n = rows * cols;
  OFL_REQUIRE(n_seeds >= 0 && n_seeds <= n && n_heights >= 0 && n_heights <= n, OFL_ERR_INVALID,
              "seed / flat_height count out of range");
  FlatsWork w;
  rc = carve(workspace, workspace_bytes, n, &w);
  if (rc != OFL_OK) return rc;
  PhaseScope ps(PHASE_FLATS, st);
  OFL_CUDA(cudaMemsetAsync(w.cnt, 0, CNT_SLOTS * sizeof(unsigned), st));
  const unsigned ns = (unsigned)n_seeds;
  OFL_CUDA(cudaMemcpyAsync(w.cnt + CNT_SEEDS, &ns, sizeof(unsigned), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaStreamSynchronize(st));  // `ns` is a stack variable
  if (n_seeds) OFL_CUDA(cudaMemcpyAsync(w.q1, seeds, (size_t)n_seeds * sizeof(int), cudaMemcpyDeviceToDevice, st));
  int* acc = w.parent;  // the sweep's own maxima; merged below so untouched labels keep the caller's value
  OFL_CUDA(cudaMemsetAsync(acc, 0, (size_t)n * sizeof(int), st));
  rc = run_gradient((int)rows, (int)cols, labels, fdr, towards ? SWEEP_TOWARDS_NEGATED : SWEEP_AWAY, false, flat_mask,
                    flat_height, acc, w, nullptr, st);
  if (rc != OFL_OK) return rc;
  if (!towards && n_heights) {
    flat_height_merge_kernel<<<blocks_for(n_heights, FL_THREADS), FL_THREADS, 0, st>>>((int)n_heights, acc, flat_height);
    OFL_CHECK_LAUNCH();
  }
  return OFL_OK;
}

int launch_masked_flow_dirs(const int* flat_mask, const int* labels, uint8_t* fdr, int64_t rows, int64_t cols,
                            cudaStream_t st) {
  int rc = check_shape(rows, cols);
  if (rc != OFL_OK) return rc;
  PhaseScope ps(PHASE_FLATS, st);
  const int aligned = (reinterpret_cast<uintptr_t>(fdr) & 3u) == 0;
  flat_masked_dirs_kernel<<<blocks_for((rows * cols + 3) / 4, FL_THREADS), FL_THREADS, 0, st>>>(flat_mask, labels, fdr,
                                                                                                  (int)rows, (int)cols, aligned);
  OFL_CHECK_LAUNCH();
  return OFL_OK;
}

}  // namespace ofl
