// accumulation.cu -- D8 flow accumulation (upstream cell counts) for sm_100a.
//
// Replaces single_tile_flow_accumulation / get_next_cell / follow_path
// (reference src/overflow/flow_accumulation.py:13-158; Barnes 2016, arXiv 1608.04431,
// Alg. 1 and 2, the paper cited at flow_accumulation.py:61,100).
//
// The reference walks a FIFO over the whole raster.  On a B200 the raster is far larger
// than on-chip memory and a global-memory queue would move >40 B/cell with atomics, so the
// sweep is organised the way the cited paper organises its tiles, with a CTA's shared memory
// as the "tile":
//
//   pass A  (acc_tile_kernel<false>)  one CTA per 64x64 tile: TMA-load the codes plus a
//           one-cell halo, accumulate every flow path that stays inside the tile (frontier
//           propagation in shared memory: missing-upstream counts, one shared atomic per edge,
//           warp-ballot compacted frontier queues), follow each perimeter cell
//           to where its path leaves the tile (Alg. 2), and emit the reduced graph: for each
//           of the tile's <=252 perimeter cells its successor perimeter cell in the next
//           tile, plus the locally accumulated counts that cross the tile edge.
//   solve   (pj_round_kernel)         the reduced graph is a forest over ~6% of the cells;
//           subtree sums over it by pointer doubling: O(log depth) rounds of
//           "add my sum to my 2^j-th ancestor, then jump", integer atomics, exact.
//   pass B  (acc_tile_kernel<true>)   re-run the tile relaxation seeded with each perimeter
//           cell's inflow from outside the tile and write the final int64 counts.
//
// HBM traffic per cell: 1 B (codes) in pass A, 1 B + 8 B in pass B, plus ~1.5 B of reduced
// graph -- against 9 B/cell compulsory.  Long drainage chains cost O(log) rounds on the
// reduced graph instead of O(length) sweeps.
//
// Edge rule (flow_accumulation.py:116-124): u -> c is an edge iff code(u) in 0..7, c lies
// inside the raster and code(c) != 9.  Codes >= 8 have no downstream cell (the reference's
// out-of-bounds NEIGHBOR_OFFSETS read); NODATA cells end at -9998 (:119-121,129-137).
#include <type_traits>

#include "common.cuh"

namespace ofl {

constexpr int AT = 64;                 // tile side (cells)
constexpr int AT_SHIFT = 6;
constexpr int ACS_W = 96;              // code tile pitch: columns x0-16 .. x0+79 (TMA box inner = 96 B, 16-B aligned start)
constexpr int ACS_H = AT + 2;          // rows y0-1 .. y0+64
constexpr int ACS_X0 = 16;
constexpr int ACS_Y0 = 1;
constexpr uint32_t ACS_BYTES = ACS_W * ACS_H;
constexpr int SLOTS = 4 * AT;          // perimeter slots per tile (top, bottom, left, right)
constexpr int ACC_THREADS = 256;
constexpr uint8_t CODE_OUTSIDE = 0xFF; // halo / partial-tile positions outside the raster

// node kinds, stored in link[] bits 8..9
constexpr uint16_t KIND_TERM = 0;        // path ends inside the raster (pit, or downstream is NODATA)
constexpr uint16_t KIND_TILE_EXIT = 1;   // path continues in the next tile (succ >= 0)
constexpr uint16_t KIND_RASTER_EXIT = 2; // path leaves the raster at the link cell

__device__ __forceinline__ int dir_dy(int code) { return ((0xA901 >> (2 * code)) & 3) - 1; }  // dy+1 = 1,0,0,0,1,2,2,2
__device__ __forceinline__ int dir_dx(int code) { return ((0x901A >> (2 * code)) & 3) - 1; }  // dx+1 = 2,2,1,0,0,0,1,2

__device__ __forceinline__ int slot_of(int y, int x, int h, int w) {
  if (y == 0) return x;
  if (y == h - 1) return AT + x;
  if (x == 0) return 2 * AT + y;
  if (x == w - 1) return 3 * AT + y;
  return -1;
}

__device__ __forceinline__ bool cell_of_slot(int s, int h, int w, int& y, int& x) {
  const int side = s >> AT_SHIFT, k = s & (AT - 1);
  if (side == 0) {
    y = 0;
    x = k;
    return k < w;
  }
  if (side == 1) {
    y = h - 1;
    x = k;
    return k < w && h > 1;
  }
  if (side == 2) {
    y = k;
    x = 0;
    return k > 0 && k < h - 1;
  }
  y = k;
  x = w - 1;
  return k > 0 && k < h - 1 && w > 1;
}

struct AccParams {
  int rows, cols;  // raster size
  int ntx, nty;    // tiles per row / column
  int32_t* succ;   // [ntiles*SLOTS] successor node or -1
  uint16_t* link;  // [ntiles*SLOTS] exit slot | kind << 8
  unsigned long long* S;  // [ntiles*SLOTS] pass A: base inflow; pass B: total inflow from outside the tile
  long long* fac;
  int64_t ld_fac;
  int* err;  // device flag: set to 1 when a cycle is detected
};

__device__ __forceinline__ int node_of_cell(int gy, int gx, const AccParams& p) {
  const int ty = gy >> AT_SHIFT, tx = gx >> AT_SHIFT;
  const int h = min(AT, p.rows - (ty << AT_SHIFT)), w = min(AT, p.cols - (tx << AT_SHIFT));
  return (ty * p.ntx + tx) * SLOTS + slot_of(gy & (AT - 1), gx & (AT - 1), h, w);
}

// ---------------------------------------------------------------- in-tile frontier propagation
// Kahn's algorithm inside one 64x64 tile, frontier kept in per-warp queues in shared memory:
//   * every cell carries the number of in-tile upstream neighbours still missing;
//   * a finished cell hands its count to its downstream cell with ONE shared-memory atomic that
//     also decrements the downstream's missing-count; the thread that brings it to zero appends the
//     downstream cell to its warp's queue (ballot + popc compaction, no queue atomics);
//   * each warp pops 32 cells per trip until its own queue is dry.  A warp's queue only grows while
//     it scans its 512 source candidates, so 512 entries always suffice.
// Pass A packs [missing:4 | count:28] in one 32-bit word, so the hand-off atomic carries the value.
// Pass B needs 64-bit counts (seeds from outside the tile): missing-counts are packed four per
// 32-bit word, values live in a 64-bit array and are pulled from the upstream neighbours when a cell
// is popped (its missing-count reaching zero orders those stores before the pull).
constexpr int QCAP = 512;

struct TileLut {
  int off[8];  // cell-index offset of the downstream neighbour per direction code
};

__device__ __forceinline__ void tile_lut_init(int* off_s) {
  if (threadIdx.x < 8) off_s[threadIdx.x] = dir_dy(threadIdx.x) * AT + dir_dx(threadIdx.x);
}

template <bool FINAL>
struct TileSmem {
  static constexpr int CS = 0;
  static constexpr int VAL = 6400;  // ACS_BYTES rounded up to 128
  static constexpr int WORD = VAL + (FINAL ? AT * AT * 8 : 0);
  static constexpr int DN = WORD + (FINAL ? AT * AT : AT * AT * 4);
  static constexpr int UPM = DN + AT * AT;
  static constexpr int Q = UPM + (FINAL ? AT * AT : 0);
  static constexpr int OFF = Q + (ACC_THREADS / 32) * QCAP * 2;
  static constexpr int BAR = OFF + 32;
  static constexpr int BYTES = BAR + 16;
};
static_assert(ACS_BYTES <= 6400, "code tile does not fit its shared-memory slot");

template <bool FINAL>
__global__ void __launch_bounds__(ACC_THREADS) acc_tile_kernel(const __grid_constant__ CUtensorMap tm,
                                                                const AccParams p) {
  // dynamic shared memory, carved by TileSmem<FINAL>
  extern __shared__ __align__(128) uint8_t smem_raw[];
  using SM = TileSmem<FINAL>;
  uint8_t* cs = smem_raw + SM::CS;                                                  // codes + halo (TMA destination)
  unsigned long long* val64 = reinterpret_cast<unsigned long long*>(smem_raw + SM::VAL);  // pass B: 64-bit counts
  uint32_t* word = reinterpret_cast<uint32_t*>(smem_raw + SM::WORD);  // A: [missing:4|count:28]; B: 4 counts/word
  uint8_t* dn = smem_raw + SM::DN;                                                  // in-tile downstream direction, 8 = none
  uint8_t* upm = smem_raw + SM::UPM;                                                // pass B: upstream-neighbour mask
  uint16_t(*q)[QCAP] = reinterpret_cast<uint16_t(*)[QCAP]>(smem_raw + SM::Q);
  int* off_s = reinterpret_cast<int*>(smem_raw + SM::OFF);
  uint64_t& bar = *reinterpret_cast<uint64_t*>(smem_raw + SM::BAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int ty = tile / p.ntx, tx = tile - ty * p.ntx;
  const int y0 = ty << AT_SHIFT, x0 = tx << AT_SHIFT;
  const int h = min(AT, p.rows - y0), w = min(AT, p.cols - x0);

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
    mbar_arrive_expect_tx(&bar, ACS_BYTES);
    tma_load_2d(cs, &tm, x0 - ACS_X0, y0 - ACS_Y0, &bar);
  }
  tile_lut_init(off_s);
  __syncthreads();
  mbar_wait(&bar, 0);

  // positions outside the raster were zero-filled by TMA (code 0 = east): mark them
  const bool edge_tile = (y0 == 0) || (x0 == 0) || (y0 + AT + 1 > p.rows) || (x0 + AT + 1 > p.cols);
  if (edge_tile) {
    for (int idx = tid; idx < (int)ACS_BYTES; idx += ACC_THREADS) {
      const int yy = idx / ACS_W, xx = idx - yy * ACS_W;
      const int gy = y0 - ACS_Y0 + yy, gx = x0 - ACS_X0 + xx;
      if (gy < 0 || gy >= p.rows || gx < 0 || gx >= p.cols) cs[idx] = CODE_OUTSIDE;
    }
    __syncthreads();
  }

  // pass B keeps four 8-bit missing-counts per word; cell c -> word c>>2, byte c&3
  uint32_t* cnt4 = word;

  // ---- phase 1: missing-counts, downstream direction, seeds.  Thread owns rows 8*warp..+7 of
  //      columns lane and lane+32 (lanes touch consecutive shared-memory words).
  uint32_t srcmask = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int x = lane + 32 * half;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int y = 8 * warp + k;
      const int cell = y * AT + x;
      const uint8_t* c = cs + (y + ACS_Y0) * ACS_W + (x + ACS_X0);
      const uint32_t own = c[0];
      uint32_t up = 0, dnb = 8;
      if (own != OFL_DIR_NODATA && own != CODE_OUTSIDE) {
        // neighbour in direction i flows into this cell iff its code is the opposite direction
        up |= (x + 1 < w && c[1] == 4) ? 1u : 0u;
        up |= (y > 0 && x + 1 < w && c[-ACS_W + 1] == 5) ? 2u : 0u;
        up |= (y > 0 && c[-ACS_W] == 6) ? 4u : 0u;
        up |= (y > 0 && x > 0 && c[-ACS_W - 1] == 7) ? 8u : 0u;
        up |= (x > 0 && c[-1] == 0) ? 16u : 0u;
        up |= (y + 1 < h && x > 0 && c[ACS_W - 1] == 1) ? 32u : 0u;
        up |= (y + 1 < h && c[ACS_W] == 2) ? 64u : 0u;
        up |= (y + 1 < h && x + 1 < w && c[ACS_W + 1] == 3) ? 128u : 0u;
        if (own < 8) {
          const int ny = y + dir_dy(own), nx = x + dir_dx(own);
          if (ny >= 0 && ny < h && nx >= 0 && nx < w && c[dir_dy(own) * ACS_W + dir_dx(own)] != OFL_DIR_NODATA)
            dnb = own;
        }
      }
      const uint32_t missing = __popc(up);
      dn[cell] = (uint8_t)dnb;
      if (FINAL) {
        unsigned long long seed = 0;
        if (own != CODE_OUTSIDE) {
          const int s = slot_of(y, x, h, w);
          if (s >= 0) seed = p.S[(size_t)tile * SLOTS + s];
        }
        val64[cell] = seed;
        upm[cell] = (uint8_t)up;
        // four consecutive cells of a row belong to four consecutive lanes: assemble the count word
        uint32_t packed = missing << (8 * (x & 3));
        packed |= __shfl_xor_sync(0xffffffffu, packed, 1);
        packed |= __shfl_xor_sync(0xffffffffu, packed, 2);
        if ((x & 3) == 0) cnt4[cell >> 2] = packed;
      } else {
        word[cell] = missing << 28;
      }
      if (missing == 0) srcmask |= 1u << (k + 8 * half);
    }
  }
  __syncthreads();

  // ---- phase 2 + 3: sources first (one per lane per step), then drain this warp's queue
  uint16_t* myq = q[warp];
  uint32_t head = 0, tail = 0;
  auto process = [&](int cell, bool active) {
    // finish `cell`, hand its count downstream, return the downstream cell if this made it ready
    bool ready = false;
    int nxt = 0;
    if (active) {
      const uint32_t d = dn[cell];
      if (FINAL) {
        unsigned long long v = val64[cell] + 1;
        uint32_t um = upm[cell];
        while (um) {
          const int i = __ffs(um) - 1;
          um &= um - 1;
          v += val64[cell + off_s[i]];
        }
        val64[cell] = v;
        if (d < 8) {
          nxt = cell + off_s[d];
          __threadfence_block();  // publish val64[cell] before the count that releases the downstream cell
          const uint32_t sh = 8 * (nxt & 3);
          const uint32_t old = atomicSub(&cnt4[nxt >> 2], 1u << sh);
          ready = ((old >> sh) & 0xFF) == 1;
          if (ready) __threadfence_block();
        }
      } else {
        const uint32_t v = (word[cell] & 0x0FFFFFFFu) + 1;
        word[cell] = v;
        if (d < 8) {
          nxt = cell + off_s[d];
          const uint32_t old = atomicAdd(&word[nxt], v - (1u << 28));
          ready = (old >> 28) == 1;
        }
      }
    }
    const uint32_t bal = __ballot_sync(0xffffffffu, ready);
    if (ready) myq[(tail + __popc(bal & ((1u << lane) - 1))) & (QCAP - 1)] = (uint16_t)nxt;
    tail += __popc(bal);
  };

#pragma unroll 1
  for (int j = 0; j < 16; ++j) {
    const int cell = (8 * warp + (j & 7)) * AT + lane + 32 * (j >> 3);
    process(cell, (srcmask >> j) & 1);
  }
  __syncwarp();
  while (head != tail) {
    const uint32_t n = tail - head;
    const uint32_t take = n < 32 ? n : 32;
    const bool active = lane < take;
    const int cell = active ? myq[(head + lane) & (QCAP - 1)] : 0;
    head += take;
    __syncwarp();
    process(cell, active);
    __syncwarp();
  }
  __syncthreads();

  if (!FINAL) {
    // ---- Alg. 2: follow every perimeter cell to where its path leaves the tile; emit the reduced graph
    const int s = tid;  // SLOTS == ACC_THREADS
    int y, x;
    int32_t succ = -1;
    uint16_t lk = KIND_TERM << 8;
    if (cell_of_slot(s, h, w, y, x) && cs[(y + ACS_Y0) * ACS_W + x + ACS_X0] != CODE_OUTSIDE) {
      int cur = y * AT + x;
      for (int steps = 0;; ++steps) {
        const uint32_t d = dn[cur];
        if (d >= 8) break;
        cur += off_s[d];
        if (steps > AT * AT) {
          atomicExch(p.err, 1);
          break;
        }
      }
      // classify the end of the in-tile path
      const int cy = cur >> AT_SHIFT, cx = cur & (AT - 1);
      uint16_t kind = KIND_TERM;
      const int code = cs[(cy + ACS_Y0) * ACS_W + cx + ACS_X0];
      if (code < 8) {
        const int ny = cy + dir_dy(code), nx = cx + dir_dx(code);
        const int dcode = cs[(ny + ACS_Y0) * ACS_W + nx + ACS_X0];
        if (dcode == CODE_OUTSIDE) {
          kind = KIND_RASTER_EXIT;
        } else if (dcode != OFL_DIR_NODATA) {
          kind = KIND_TILE_EXIT;  // dn == 8 with a live downstream cell: it lies in the next tile
          succ = node_of_cell(y0 + ny, x0 + nx, p);
        }
      }
      const int ls = slot_of(cy, cx, h, w);
      lk = (uint16_t)((ls < 0 ? 0 : ls) | (kind << 8));
      // this cell's own edge across the tile boundary carries its local count to the next tile
      const int own = cs[(y + ACS_Y0) * ACS_W + x + ACS_X0];
      if (own < 8 && dn[y * AT + x] >= 8) {
        const int ny = y + dir_dy(own), nx = x + dir_dx(own);
        const int dcode = cs[(ny + ACS_Y0) * ACS_W + nx + ACS_X0];
        if (dcode != CODE_OUTSIDE && dcode != OFL_DIR_NODATA) {
          const uint32_t wv = word[y * AT + x];
          if (wv >> 28) atomicExch(p.err, 1);  // never finished: the tile holds a cycle
          atomicAdd(&p.S[node_of_cell(y0 + ny, x0 + nx, p)], (unsigned long long)(wv & 0x0FFFFFFFu));
        }
      }
    }
    p.succ[(size_t)tile * SLOTS + s] = succ;
    p.link[(size_t)tile * SLOTS + s] = lk;
  } else {
    // ---- final counts: lane-contiguous int64 stores (256 B per half row); NODATA cells get -9998
    bool stuck = false;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int x = lane + 32 * half;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int y = 8 * warp + k;
        if (y < h && x < w) {
          const int cell = y * AT + x;
          const uint8_t own = cs[(y + ACS_Y0) * ACS_W + x + ACS_X0];
          stuck |= ((cnt4[cell >> 2] >> (8 * (cell & 3))) & 0xFF) != 0;
          const long long v = (own == OFL_DIR_NODATA) ? (long long)OFL_FAC_NODATA_EMITTED : (long long)val64[cell];
          p.fac[(int64_t)(y0 + y) * p.ld_fac + (x0 + x)] = v;
        }
      }
    }
    if (stuck) atomicExch(p.err, 1);  // a missing-count never reached zero: cycle
  }
}

// ---------------------------------------------------------------- reduced-graph solve
// ptr[u] >= 0: u's 2^j-th ancestor.  ptr[u] < 0: ~root(u), the chain is exhausted.
__global__ void pj_init_kernel(const int32_t* __restrict__ succ, int32_t* __restrict__ ptr, int64_t n) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    const int32_t s = succ[u];
    ptr[u] = s >= 0 ? s : ~(int32_t)u;
  }
}

// Round j: S_{j+1}[v] = S_j[v] + sum of S_j[w] over nodes w whose 2^j-th ancestor is v.
// Deltas land in d_next and are folded into S by their owner at the start of the next round.
__global__ void pj_round_kernel(const int32_t* __restrict__ ptr_in, int32_t* __restrict__ ptr_out,
                                unsigned long long* __restrict__ S, unsigned long long* __restrict__ d_prev,
                                unsigned long long* __restrict__ d_next, const int* __restrict__ active_in,
                                int* __restrict__ active_out, int64_t n) {
  if (*active_in == 0) return;  // converged in an earlier round
  bool any = false;
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    unsigned long long s = S[u];
    const unsigned long long dp = d_prev[u];
    if (dp) {
      s += dp;
      S[u] = s;
      d_prev[u] = 0;
    }
    const int32_t a = ptr_in[u];
    int32_t q = a;
    if (a >= 0) {
      if (s) atomicAdd(&d_next[a], s);
      q = ptr_in[a];
      any |= (q >= 0);
    }
    ptr_out[u] = q;
  }
  if (__any_sync(0xffffffffu, any) && (threadIdx.x & 31) == 0) *active_out = 1;
}

__global__ void pj_fold_kernel(unsigned long long* __restrict__ S, const unsigned long long* __restrict__ d0,
                               const unsigned long long* __restrict__ d1, int64_t n) {
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < n; u += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long d = d0[u] + d1[u];
    if (d) S[u] += d;
  }
}

// ---------------------------------------------------------------- links for the raster perimeter
// perimeter_indices order (flow_accumulation.py:40-51): every row's left then right cell, then for
// columns 1..C-2 the top then the bottom cell.
__global__ void links_kernel(const uint8_t* __restrict__ fdr, int64_t ld_fdr, const int32_t* __restrict__ root_ptr,
                             const uint16_t* __restrict__ link, AccParams p, long long* __restrict__ out, int64_t n) {
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    int r, c;
    if (k < 2 * (int64_t)p.rows) {
      r = (int)(k >> 1);
      c = (k & 1) ? p.cols - 1 : 0;
    } else {
      const int64_t kk = k - 2 * (int64_t)p.rows;
      c = 1 + (int)(kk >> 1);
      r = (kk & 1) ? p.rows - 1 : 0;
    }
    long long lr, lc;
    const int code = fdr[(int64_t)r * ld_fdr + c];
    if (code >= 8) {
      // reference quirk: the out-of-bounds offset read sends codes 8/9 "outside" on the first step
      lr = lc = OFL_LINK_EXTERNAL;
    } else {
      const int u = node_of_cell(r, c, p);
      const int32_t pr = root_ptr[u];
      const int root = pr < 0 ? ~pr : u;
      const uint16_t lk = link[root];
      const int kind = lk >> 8;
      if (kind == KIND_RASTER_EXIT) {
        const int tile = root / SLOTS;
        const int ty = tile / p.ntx, tx = tile - ty * p.ntx;
        const int h = min(AT, p.rows - (ty << AT_SHIFT)), w = min(AT, p.cols - (tx << AT_SHIFT));
        int y, x;
        cell_of_slot(lk & 0xFF, h, w, y, x);
        const int er = (ty << AT_SHIFT) + y, ec = (tx << AT_SHIFT) + x;
        if (er == r && ec == c) {
          lr = lc = OFL_LINK_EXTERNAL;
        } else {
          lr = er;
          lc = ec;
        }
      } else {
        lr = lc = OFL_LINK_TERMINATES;
      }
    }
    out[2 * k] = lr;
    out[2 * k + 1] = lc;
  }
}

// ---------------------------------------------------------------- recurrence checker
__global__ void check_kernel(const uint8_t* __restrict__ fdr, int64_t ld_fdr, const long long* __restrict__ fac,
                             int64_t ld_fac, int rows, int cols, unsigned long long* __restrict__ n_bad) {
  const int64_t n = (int64_t)rows * cols;
  unsigned long long bad = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const int code = fdr[(int64_t)r * ld_fdr + c];
    long long want;
    if (code == OFL_DIR_NODATA) {
      want = OFL_FAC_NODATA_EMITTED;
    } else {
      want = 1;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const int ur = r + dir_dy(d), uc = c + dir_dx(d);
        if (ur < 0 || ur >= rows || uc < 0 || uc >= cols) continue;
        if (fdr[(int64_t)ur * ld_fdr + uc] == ((d + 4) & 7)) want += fac[(int64_t)ur * ld_fac + uc];
      }
    }
    bad += (fac[(int64_t)r * ld_fac + c] != want);
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_down_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, bad);
}

// ---------------------------------------------------------------- host orchestration (device pointers)
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int64_t perimeter_count(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 0;
  const int64_t inner = cols - 2 > 0 ? cols - 2 : 0;
  return 2 * rows + 2 * inner;
}

struct AccLayout {
  int64_t n_nodes;
  size_t off_succ, off_ptr, off_S, off_d0, off_d1, off_link, off_flags, total;
};

static AccLayout acc_layout(int64_t rows, int64_t cols) {
  AccLayout L;
  const int64_t nty = (rows + AT - 1) / AT, ntx = (cols + AT - 1) / AT;
  L.n_nodes = nty * ntx * SLOTS;
  size_t o = 0;
  L.off_succ = o;
  o = align_up(o + (size_t)L.n_nodes * 4, 256);
  L.off_ptr = o;
  o = align_up(o + (size_t)L.n_nodes * 4, 256);
  L.off_S = o;
  o = align_up(o + (size_t)L.n_nodes * 8, 256);
  L.off_d0 = o;
  o = align_up(o + (size_t)L.n_nodes * 8, 256);
  L.off_d1 = o;
  o = align_up(o + (size_t)L.n_nodes * 8, 256);
  L.off_link = o;
  o = align_up(o + (size_t)L.n_nodes * 2, 256);
  L.off_flags = o;
  o = align_up(o + 64 * sizeof(int), 256);
  L.total = o;
  return L;
}

size_t accumulation_workspace_bytes(int64_t rows, int64_t cols) {
  if (rows <= 0 || cols <= 0) return 256;
  return acc_layout(rows, cols).total;
}

constexpr int PJ_MAX_ROUNDS = 40;  // flags[0..40]: active-before-round j; flags[48]: cycle error

int launch_accumulation(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, long long* fac,
                        int64_t ld_fac, long long* perim_links_dev, void* workspace, size_t workspace_bytes,
                        cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return OFL_OK;
  OFL_REQUIRE(rows < (1ll << 30) && cols < (1ll << 30), OFL_ERR_INVALID, "raster dimension too large");
  OFL_REQUIRE((reinterpret_cast<uintptr_t>(fdr) & 15) == 0 && (ld_fdr % 16) == 0 && ld_fdr >= cols, OFL_ERR_ALIGNMENT,
              "fdr must be 16-byte aligned with ld_fdr %% 16 == 0 (ld_fdr=%lld)", (long long)ld_fdr);
  OFL_REQUIRE((reinterpret_cast<uintptr_t>(fac) & 7) == 0 && ld_fac >= cols, OFL_ERR_ALIGNMENT,
              "fac must be 8-byte aligned");
  const AccLayout L = acc_layout(rows, cols);
  OFL_REQUIRE(L.n_nodes < (1ll << 31), OFL_ERR_INVALID, "raster too large for 32-bit perimeter node ids");
  OFL_REQUIRE(workspace != nullptr && workspace_bytes >= L.total, OFL_ERR_WORKSPACE,
              "accumulation workspace too small: need %zu bytes, have %zu", L.total, workspace_bytes);
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  AccParams p;
  p.rows = (int)rows;
  p.cols = (int)cols;
  p.nty = (int)((rows + AT - 1) / AT);
  p.ntx = (int)((cols + AT - 1) / AT);
  p.succ = reinterpret_cast<int32_t*>(ws + L.off_succ);
  int32_t* ptr_b = reinterpret_cast<int32_t*>(ws + L.off_ptr);
  p.link = reinterpret_cast<uint16_t*>(ws + L.off_link);
  p.S = reinterpret_cast<unsigned long long*>(ws + L.off_S);
  unsigned long long* d0 = reinterpret_cast<unsigned long long*>(ws + L.off_d0);
  unsigned long long* d1 = reinterpret_cast<unsigned long long*>(ws + L.off_d1);
  int* flags = reinterpret_cast<int*>(ws + L.off_flags);
  p.fac = fac;
  p.ld_fac = ld_fac;
  p.err = flags + 48;
  const int64_t ntiles = (int64_t)p.nty * p.ntx;
  OFL_REQUIRE(ntiles < (1ll << 31), OFL_ERR_INVALID, "too many tiles");

  CUtensorMap tm;
  int rc = make_tensor_map_2d(&tm, fdr, 1, (uint64_t)cols, (uint64_t)rows, (uint64_t)ld_fdr, ACS_W, ACS_H);
  if (rc != OFL_OK) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    OFL_CUDA(cudaFuncSetAttribute(acc_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TileSmem<false>::BYTES));
    OFL_CUDA(cudaFuncSetAttribute(acc_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TileSmem<true>::BYTES));
    attr_set = true;
  }

  // S, d0, d1 are contiguous: one memset; flags: active-before-round-0 = 1, the rest 0
  OFL_CUDA(cudaMemsetAsync(ws + L.off_S, 0, L.off_link - L.off_S, st));
  OFL_CUDA(cudaMemsetAsync(flags, 0, 64 * sizeof(int), st));

  {
    PhaseScope ps(PHASE_ACC_TILE_A, st);
    acc_tile_kernel<false><<<(unsigned)ntiles, ACC_THREADS, TileSmem<false>::BYTES, st>>>(tm, p);
  }
  OFL_CHECK_LAUNCH();

  // reduced-graph solve
  const int sms = sm_count();
  const int pj_blocks = (int)(((L.n_nodes + 255) / 256) < (int64_t)sms * 8 ? ((L.n_nodes + 255) / 256) : (int64_t)sms * 8);
  PhaseScope* solve_scope = new PhaseScope(PHASE_ACC_SOLVE, st);
  int32_t* ptr_cur = ptr_b;
  int32_t* ptr_nxt = p.succ;  // succ is dead once pj_init has consumed it
  pj_init_kernel<<<pj_blocks, 256, 0, st>>>(p.succ, ptr_cur, L.n_nodes);
  OFL_CHECK_LAUNCH();
  {
    const int one = 1;
    OFL_CUDA(cudaMemcpyAsync(flags, &one, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  int max_rounds = 2;
  while ((1ll << (max_rounds - 1)) < L.n_nodes && max_rounds < PJ_MAX_ROUNDS) ++max_rounds;
  for (int j = 0; j < max_rounds; ++j) {
    pj_round_kernel<<<pj_blocks, 256, 0, st>>>(ptr_cur, ptr_nxt, p.S, (j & 1) ? d1 : d0, (j & 1) ? d0 : d1,
                                               flags + j, flags + j + 1, L.n_nodes);
    OFL_CHECK_LAUNCH();
    int32_t* t = ptr_cur;
    ptr_cur = ptr_nxt;
    ptr_nxt = t;
  }
  pj_fold_kernel<<<pj_blocks, 256, 0, st>>>(p.S, d0, d1, L.n_nodes);
  delete solve_scope;
  OFL_CHECK_LAUNCH();

  {
    PhaseScope ps(PHASE_ACC_TILE_B, st);
    acc_tile_kernel<true><<<(unsigned)ntiles, ACC_THREADS, TileSmem<true>::BYTES, st>>>(tm, p);
  }
  OFL_CHECK_LAUNCH();

  // Rounds after convergence return early without writing ptr_out, so the converged pointers sit in
  // whichever buffer the last ACTIVE round wrote: round j writes buffer (j even ? succ : ptr_b).
  int h_flags[64];
  OFL_CUDA(cudaMemcpyAsync(h_flags, flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  OFL_REQUIRE(h_flags[48] == 0, OFL_ERR_CYCLE, "flow-direction raster contains a cycle");
  int last_active = -1;
  for (int j = 0; j < max_rounds; ++j)
    if (h_flags[j]) last_active = j;
  OFL_REQUIRE(h_flags[max_rounds] == 0, OFL_ERR_CYCLE, "flow-direction raster contains a cycle (perimeter graph)");
  if (perim_links_dev) {
    const int32_t* roots = (last_active < 0) ? ptr_b : ((last_active & 1) == 0 ? p.succ : ptr_b);
    const int64_t n = perimeter_count(rows, cols);
    const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
    {
      PhaseScope ps(PHASE_ACC_LINKS, st);
      links_kernel<<<blocks, 256, 0, st>>>(fdr, ld_fdr, roots, p.link, p, perim_links_dev, n);
    }
    OFL_CHECK_LAUNCH();
  }
  return OFL_OK;
}

int launch_check(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, const long long* fac, int64_t ld_fac,
                 unsigned long long* n_bad_dev, cudaStream_t st) {
  OFL_CUDA(cudaMemsetAsync(n_bad_dev, 0, sizeof(unsigned long long), st));
  if (rows <= 0 || cols <= 0) return OFL_OK;
  const int64_t n = rows * cols;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < (int64_t)sm_count() * 16 ? want : (int64_t)sm_count() * 16);
  check_kernel<<<blocks, 256, 0, st>>>(fdr, ld_fdr, fac, ld_fac, (int)rows, (int)cols, n_bad_dev);
  OFL_CHECK_LAUNCH();
  return OFL_OK;
}

}  // namespace ofl
