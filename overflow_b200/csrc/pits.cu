// pits.cu -- single-cell pit breaching on the device: the reference's
// src/overflow/breach_single_cell_pits.py:9-63 (breach_single_cell_pits_in_chunk).
//
// The reference finds the pits of the incoming chunk in a parallel pass (a cell whose eight neighbours are all
// strictly higher and none of them NODATA), then visits them ONE BY ONE in row-major order: a pit looks at the
// sixteen cells two steps away and, for each that is not higher (or is NODATA), writes the mean of the two
// elevations into the cell between them.  That second pass is order dependent -- a pit may read a cell an
// earlier pit has lowered, and two pits may write the same cell, the later one winning -- but only locally: a
// pit writes within one cell of itself and reads within two, so two pits interact only when they are at most
// three cells apart (chessboard distance).  Here the pits are collected into a list and breached in rounds: a
// pit is ready when no EARLIER pit within distance three is still waiting; the ready pits of a round are more
// than three cells apart from each other, so they run concurrently on disjoint cells and every pit sees exactly
// the chunk the reference's sequential loop would show it.  The earliest waiting pit is always ready, so the
// rounds end; on terrain they number a handful.
//
// Arithmetic as numba types it for a float32 chunk: float32 sum, float64 division by two, float32 store; the
// NODATA test compares in float64.  Bit-exact against the reference (tests/golden/breach_pits.npz).
#include <math_constants.h>

#include "common.cuh"

namespace ofl {
namespace {

constexpr int PT = 256;
enum { PC_PITS = 0, PC_LEFT, PC_UNSOLVED, PC_SLOTS = 16 };

__constant__ int p_dx[8] = {1, 1, 1, 0, -1, -1, -1, 0};  // breach_single_cell_pits.py:27-31
__constant__ int p_dy[8] = {-1, 0, 1, 1, 1, 0, -1, -1};
__constant__ int p_dx2[16] = {2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1};
__constant__ int p_dy2[16] = {-2, -1, 0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2};
__constant__ int p_breach[16] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 0};

// pass 1 (:38-50): pits of the chunk as it came in -> unsolved = waiting = 1, index appended to the list
__global__ void __launch_bounds__(PT)
pits_detect_kernel(const float* __restrict__ chunk, int rows, int cols, int64_t ld, float nd_f, bool nd_exact,
                   int8_t* unsolved, uint8_t* waiting, int* list, unsigned* cnt) {
  __shared__ unsigned warp_off[PT / 32];
  __shared__ unsigned cta_base;
  const unsigned n = (unsigned)rows * (unsigned)cols;
  const unsigned i = blockIdx.x * PT + threadIdx.x;
  bool pit = false;
  if (i < n) {
    const int r = (int)(i / (unsigned)cols), c = (int)(i - (unsigned)r * (unsigned)cols);
    if (r >= 2 && r < rows - 2 && c >= 2 && c < cols - 2) {
      const float* at = chunk + (int64_t)r * ld + c;
      const float z = *at;
      // Branch-free form of :40-50.  (double)x == nodata <=> x == (float)nodata when float32 holds nodata exactly,
      // and never otherwise.  fminf skips NaN neighbours exactly as "zn <= z" is false for them; starting from NaN
      // and testing "not (lowest <= z)" keeps the reference's answer when z or every neighbour is NaN.
      float lowest = CUDART_NAN_F;
      bool nodata_seen = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float zn = at[(int64_t)p_dy[k] * ld + p_dx[k]];
        lowest = fminf(lowest, zn);
        nodata_seen = nodata_seen || zn == nd_f;
      }
      pit = !(nd_exact && (nodata_seen || z == nd_f)) && !(lowest <= z);
    }
    unsolved[i] = pit ? 1 : 0;
    waiting[i] = pit ? 1 : 0;
  }
  // one list atomic per CTA
  const unsigned m = __ballot_sync(0xffffffffu, pit);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) warp_off[w] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned run = 0;
    for (int k = 0; k < PT / 32; ++k) {
      const unsigned v = warp_off[k];
      warp_off[k] = run;
      run += v;
    }
    cta_base = run ? atomicAdd(&cnt[PC_PITS], run) : 0u;
  }
  __syncthreads();
  if (pit) list[cta_base + warp_off[w] + __popc(m & ((1u << lane) - 1u))] = (int)i;
}

// a waiting pit is ready when no earlier (row-major) pit within chessboard distance 3 is still waiting
__global__ void __launch_bounds__(PT)
pits_ready_kernel(const int* __restrict__ list, unsigned n_pits, const uint8_t* __restrict__ waiting, uint8_t* ready,
                  int rows, int cols) {
  const unsigned t = blockIdx.x * PT + threadIdx.x;
  if (t >= n_pits) return;
  const int i = list[t];
  if (!waiting[i]) return;
  const int r = i / cols, c = i - r * cols;
  bool ok = true;
  for (int dr = -3; dr <= 0 && ok; ++dr) {
    const int rr = r + dr;
    if (rr < 0) continue;
    const int c_hi = dr < 0 ? min(cols - 1, c + 3) : c - 1;  // the pit's own row: only cells before it
    for (int cc = max(0, c - 3); cc <= c_hi; ++cc)
      if (waiting[rr * cols + cc]) {
        ok = false;
        break;
      }
  }
  ready[t] = ok ? 1 : 0;
}

// pass 2 (:52-63) for the ready pits of this round
__global__ void __launch_bounds__(PT)
pits_breach_kernel(const int* __restrict__ list, unsigned n_pits, uint8_t* waiting, uint8_t* ready, float* chunk, int rows,
                   int cols, int64_t ld, double nodata, int8_t* unsolved, unsigned* cnt) {
  const unsigned t = blockIdx.x * PT + threadIdx.x;
  if (t >= n_pits || !ready[t]) return;
  ready[t] = 0;
  const int i = list[t];
  const int r = i / cols, c = i - r * cols;
  float* at = chunk + (int64_t)r * ld + c;
  const float z = *at;
  bool solved = false;
  for (int k = 0; k < 16; ++k) {  // in order: two k share a breach cell and the later one wins
    const float zn = at[(int64_t)p_dy2[k] * ld + p_dx2[k]];
    if (zn <= z || (double)zn == nodata) {
      solved = true;
      const int b = p_breach[k];
      at[(int64_t)p_dy[b] * ld + p_dx[b]] = __double2float_rn((double)__fadd_rn(z, zn) / 2.0);
    }
  }
  if (solved)
    unsolved[i] = 0;
  else
    atomicAdd(&cnt[PC_UNSOLVED], 1u);
  waiting[i] = 0;
  atomicSub(&cnt[PC_LEFT], 1u);
}

__global__ void pits_begin_kernel(unsigned* cnt) { cnt[PC_LEFT] = cnt[PC_PITS]; }

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

size_t pits_workspace_bytes(int64_t rows, int64_t cols) {
  const size_t n = (size_t)rows * (size_t)cols;
  return align256(n * sizeof(int)) + 2 * align256(n) + align256(PC_SLOTS * sizeof(unsigned));
}

// chunk: device, rows x cols with leading dimension ld, breached in place.  unsolved: device int8, dense.
// info (host, nullable): {pits found, pits left unsolved, rounds}.  Synchronises the stream.
int launch_breach_pits(float* chunk, int64_t rows, int64_t cols, int64_t ld, double nodata, int8_t* unsolved,
                       int64_t* info, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  OFL_REQUIRE(rows > 0 && cols > 0 && rows * cols < (int64_t)INT32_MAX, OFL_ERR_INVALID,
              "pit breaching works on one chunk of fewer than 2^31 cells (got %lld x %lld)", (long long)rows, (long long)cols);
  OFL_REQUIRE(workspace_bytes >= pits_workspace_bytes(rows, cols), OFL_ERR_WORKSPACE, "pits workspace too small");
  const size_t n = (size_t)rows * (size_t)cols;
  char* p = static_cast<char*>(workspace);
  int* list = reinterpret_cast<int*>(p);
  uint8_t* waiting = reinterpret_cast<uint8_t*>(p + align256(n * sizeof(int)));
  uint8_t* ready = waiting + align256(n);
  unsigned* cnt = reinterpret_cast<unsigned*>(ready + align256(n));
  PhaseScope ps(PHASE_PITS, st);
  OFL_CUDA(cudaMemsetAsync(cnt, 0, PC_SLOTS * sizeof(unsigned), st));
  const unsigned nb = (unsigned)((n + PT - 1) / PT);
  const float nd_f = (float)nodata;
  const bool nd_exact = (double)nd_f == nodata;  // false for NaN and for values float32 cannot hold: nothing matches then
  pits_detect_kernel<<<nb, PT, 0, st>>>(chunk, (int)rows, (int)cols, ld, nd_f, nd_exact, unsolved, waiting, list, cnt);
  OFL_CHECK_LAUNCH();
  pits_begin_kernel<<<1, 1, 0, st>>>(cnt);
  OFL_CHECK_LAUNCH();
  unsigned* h = nullptr;
  int rc = pinned_get(64, reinterpret_cast<void**>(&h));
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(h, cnt, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  const unsigned n_pits = h[PC_PITS];
  unsigned left = n_pits;
  int64_t rounds = 0;
  if (n_pits) {
    OFL_CUDA(cudaMemsetAsync(ready, 0, n_pits, st));
    const unsigned nbp = (n_pits + PT - 1) / PT;
    int batch = 2;  // rounds between two looks at the counter: a finished chunk costs a few empty launches at most
    while (left) {
      for (int b = 0; b < batch; ++b) {
        pits_ready_kernel<<<nbp, PT, 0, st>>>(list, n_pits, waiting, ready, (int)rows, (int)cols);
        OFL_CHECK_LAUNCH();
        pits_breach_kernel<<<nbp, PT, 0, st>>>(list, n_pits, waiting, ready, chunk, (int)rows, (int)cols, ld, nodata,
                                               unsolved, cnt);
        OFL_CHECK_LAUNCH();
      }
      rounds += batch;
      OFL_CUDA(cudaMemcpyAsync(h, cnt, 3 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
      OFL_CUDA(cudaStreamSynchronize(st));
      OFL_REQUIRE(h[PC_LEFT] < left, OFL_ERR_INVALID, "pit breaching made no progress (%u pits left)", left);
      left = h[PC_LEFT];
      if (batch < 16) batch *= 2;
    }
  }
  if (info) {
    info[0] = n_pits;
    info[1] = n_pits ? h[PC_UNSOLVED] : 0;
    info[2] = rounds;
  }
  return OFL_OK;
}

}  // namespace ofl
