// api.cu -- the extern "C" surface declared in include/overflow_b200.h.
// Host-pointer calls stage through library-owned device buffers; device-pointer calls launch directly.
#include <cstdlib>

#include "common.cuh"

namespace ofl {
const char* last_error();
int init_device(int device);
int ensure_init();
int shutdown();
int64_t launch_count();
void launch_count_reset();
void phase_timing_enable(int on);
int phase_timing_read(double* ms, int64_t* counts, int n, int reset);

int launch_direction(const float* dem, int64_t in_rows, int64_t cols, int64_t ld_dem, double nodata, uint8_t* fdr,
                     int64_t rows, int64_t ld_fdr, int y_off, cudaStream_t st);
int launch_direction_generic(const void* dem, int kind, int64_t in_rows, int64_t cols, int64_t ld_dem, double nodata,
                             uint8_t* fdr, int64_t rows, int64_t ld_fdr, int y_off, cudaStream_t st);
int launch_fill_border(uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld, int value, cudaStream_t st);
int64_t perimeter_count(int64_t rows, int64_t cols);
size_t accumulation_workspace_bytes(int64_t rows, int64_t cols);
int launch_accumulation(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, long long* fac, int64_t ld_fac,
                        long long* perim_links_dev, void* workspace, size_t workspace_bytes, cudaStream_t st,
                        bool prepared = false, bool trusted_codes = false, const long long* perim_inflow_dev = nullptr);
int accumulation_prepare(int64_t rows, int64_t cols, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t strip_workspace_bytes(int64_t rows, int64_t cols);
size_t strip_boundary_workspace_bytes(int n_strips, int64_t cols);
size_t strip_record_bytes(int64_t cols);
int strip_accum_local(const uint8_t* fdr_halo, int64_t rows, int64_t cols, int64_t ld_fdr, int has_above, int has_below,
                      long long* fac, int64_t ld_fac, void* workspace, size_t workspace_bytes, void* record,
                      cudaStream_t st);
int strip_boundary_solve(const void* records_all, int n_strips, int64_t cols, long long* J_all, void* workspace,
                         size_t workspace_bytes, cudaStream_t st);
int strip_collect_flags(const void* strip_workspace, int64_t rows, int64_t cols, const void* boundary_workspace,
                        int n_strips, int32_t* flags_dev, cudaStream_t st);
int strip_accum_final(const uint8_t* fdr_halo, int64_t rows, int64_t cols, int64_t ld_fdr, int has_above, int has_below,
                      const long long* J_mine, void* workspace, size_t workspace_bytes, long long* fac, int64_t ld_fac,
                      cudaStream_t st);
int launch_check(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, const long long* fac, int64_t ld_fac,
                 unsigned long long* n_bad_dev, cudaStream_t st, int y_off = 0, const long long* fac_above = nullptr,
                 const long long* fac_below = nullptr);
size_t flats_workspace_bytes(int64_t rows, int64_t cols);
int launch_flat_edges(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, uint8_t* edges, int64_t* n_low,
                      int64_t* n_high, unsigned* cnt_dev, cudaStream_t st);
int launch_resolve_flats(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, int* flat_mask, int* labels,
                         int64_t* info, void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_flat_gradient(const int* labels, const uint8_t* fdr, int64_t rows, int64_t cols, const int* seeds,
                         int64_t n_seeds, int towards, int* flat_mask, int* flat_height, int64_t n_heights,
                         void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_masked_flow_dirs(const int* flat_mask, const int* labels, uint8_t* fdr, int64_t rows, int64_t cols,
                            cudaStream_t st);
size_t pits_workspace_bytes(int64_t rows, int64_t cols);
int launch_breach_pits(float* chunk, int64_t rows, int64_t cols, int64_t ld, double nodata, int8_t* unsolved,
                       int64_t* info, void* workspace, size_t workspace_bytes, cudaStream_t st);
int launch_synth(float* dem, int64_t rows, int64_t cols, int64_t ld, int64_t row0, int64_t total_rows, uint64_t seed,
                 int kind, float relief, int holes_permille, float nodata, cudaStream_t st);
}  // namespace ofl

using namespace ofl;

static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------- host rasters: banded pipeline
// A host DEM goes to the device in row bands on its own stream; the stencil of band b runs as soon as
// the first row of band b+1 has arrived, and its codes return to the host on a third stream, so the
// upload, the kernel and the download overlap (PCIe is full duplex).  The staging buffers hold the whole
// raster: accumulation needs every code before it can finish a single count.
namespace {
// Streams and events of the host pipeline belong to the device they were created on: they are keyed on the
// library's device generation and destroyed (by the cleanup hook, on their own device) when it moves on.
struct HostPipe {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaEvent_t ev_in = nullptr, ev_prep = nullptr;
  cudaEvent_t pool[2 * 256] = {};  // band events of direction_from_host, created on demand, reused by every call
  int n_pool = 0;
  int gen = -1;
};
HostPipe g_pipe;

void pipe_cleanup() {
  if (g_pipe.h2d) cudaStreamDestroy(g_pipe.h2d);
  if (g_pipe.d2h) cudaStreamDestroy(g_pipe.d2h);
  if (g_pipe.ev_in) cudaEventDestroy(g_pipe.ev_in);
  if (g_pipe.ev_prep) cudaEventDestroy(g_pipe.ev_prep);
  for (int i = 0; i < g_pipe.n_pool; ++i) cudaEventDestroy(g_pipe.pool[i]);
  g_pipe = HostPipe();
}

int pipe_streams() {
  if (g_pipe.gen == device_generation() && g_pipe.h2d) return OFL_OK;
  g_pipe = HostPipe();  // whatever an older generation held was destroyed by pipe_cleanup on its own device
  register_device_cleanup(pipe_cleanup);
  OFL_CUDA(cudaStreamCreateWithFlags(&g_pipe.h2d, cudaStreamNonBlocking));
  OFL_CUDA(cudaStreamCreateWithFlags(&g_pipe.d2h, cudaStreamNonBlocking));
  OFL_CUDA(cudaEventCreateWithFlags(&g_pipe.ev_in, cudaEventDisableTiming));
  OFL_CUDA(cudaEventCreateWithFlags(&g_pipe.ev_prep, cudaEventDisableTiming));
  g_pipe.gen = device_generation();
  return OFL_OK;
}

constexpr int64_t PIPE_BAND_BYTES = 512ll << 20;  // ~0.5 GiB of DEM per band
constexpr int PIPE_MAX_BANDS = 256;

int pipe_events(int n) {
  while (g_pipe.n_pool < n) {
    OFL_CUDA(cudaEventCreateWithFlags(&g_pipe.pool[g_pipe.n_pool], cudaEventDisableTiming));
    ++g_pipe.n_pool;
  }
  return OFL_OK;
}

// Enqueues the whole banded pipeline; returns at the first error (the caller drains the streams).
int direction_from_host_enqueue(const float* dem, int64_t in_rows, int64_t rows, int64_t cols, int64_t ld_dem,
                                double nodata, int y_off, float* d_dem, int64_t ldd, uint8_t* d_fdr, int64_t ldo,
                                uint8_t* fdr_host, int64_t ld_fdr, cudaStream_t st) {
  int64_t band_bytes = PIPE_BAND_BYTES;
  if (const char* e = getenv("OFL_PIPE_BAND_BYTES")) band_bytes = atoll(e) > 0 ? atoll(e) : band_bytes;  // tests: many small bands
  int64_t band = band_bytes / (cols * (int64_t)sizeof(float));
  band = band < 64 ? 64 : band / 64 * 64;
  if ((in_rows + band - 1) / band > PIPE_MAX_BANDS) band = ((in_rows + PIPE_MAX_BANDS - 1) / PIPE_MAX_BANDS + 63) / 64 * 64;
  const int n_in = (int)((in_rows + band - 1) / band), n_out = (int)((rows + band - 1) / band);
  int rc = pipe_events(n_in + n_out);
  if (rc != OFL_OK) return rc;
  cudaEvent_t* ev_up = g_pipe.pool;
  cudaEvent_t* ev_dir = g_pipe.pool + n_in;
  for (int b = 0; b < n_in; ++b) {
    const int64_t r0 = b * band, r1 = (r0 + band < in_rows) ? r0 + band : in_rows;
    OFL_CUDA(cudaMemcpy2DAsync(d_dem + r0 * ldd, ldd * sizeof(float), dem + r0 * ld_dem, ld_dem * sizeof(float),
                               cols * sizeof(float), r1 - r0, cudaMemcpyHostToDevice, g_pipe.h2d));
    OFL_CUDA(cudaEventRecord(ev_up[b], g_pipe.h2d));
  }
  for (int b = 0; b < n_out; ++b) {
    const int64_t r0 = b * band, r1 = (r0 + band < rows) ? r0 + band : rows;
    int64_t in_lo = r0 + y_off - 1, in_hi = r1 + y_off + 1;  // input rows [in_lo, in_hi)
    if (in_lo < 0) in_lo = 0;
    if (in_hi > in_rows) in_hi = in_rows;
    OFL_CUDA(cudaStreamWaitEvent(st, ev_up[(in_hi - 1) / band], 0));  // uploads complete in order
    rc = launch_direction(d_dem + in_lo * ldd, in_hi - in_lo, cols, ldd, nodata, d_fdr + r0 * ldo, r1 - r0, ldo,
                          (int)(r0 + y_off - in_lo), st);
    if (rc != OFL_OK) return rc;
    if (fdr_host) {
      OFL_CUDA(cudaEventRecord(ev_dir[b], st));
      OFL_CUDA(cudaStreamWaitEvent(g_pipe.d2h, ev_dir[b], 0));
      OFL_CUDA(cudaMemcpy2DAsync(fdr_host + r0 * ld_fdr, ld_fdr, d_fdr + r0 * ldo, ldo, cols, r1 - r0,
                                 cudaMemcpyDeviceToHost, g_pipe.d2h));
    }
  }
  return OFL_OK;
}

// dem: host, in_rows x cols.  d_dem / d_fdr: device staging (pitches ldd / ldo).  fdr_host may be null.
// Output row y reads input rows y + y_off - 1 .. y + y_off + 1.  With sync_downloads, or after any error, every
// stream is idle on return: the caller's host buffers are never left in flight.
int direction_from_host(const float* dem, int64_t in_rows, int64_t rows, int64_t cols, int64_t ld_dem, double nodata,
                        int y_off, float* d_dem, int64_t ldd, uint8_t* d_fdr, int64_t ldo, uint8_t* fdr_host,
                        int64_t ld_fdr, cudaStream_t st, bool sync_downloads) {
  int rc = pipe_streams();
  if (rc != OFL_OK) return rc;
  rc = direction_from_host_enqueue(dem, in_rows, rows, cols, ld_dem, nodata, y_off, d_dem, ldd, d_fdr, ldo, fdr_host,
                                   ld_fdr, st);
  if (rc != OFL_OK || sync_downloads) {
    cudaStreamSynchronize(g_pipe.h2d);
    cudaStreamSynchronize(st);
    cudaStreamSynchronize(g_pipe.d2h);
  }
  return rc;
}
}  // namespace

extern "C" {

const char* ofl_last_error(void) { return last_error(); }
int ofl_abi_version(void) { return OFL_ABI_VERSION; }
int ofl_init(int device) { return init_device(device); }
int ofl_shutdown(void) { return shutdown(); }
int64_t ofl_launch_count(void) { return launch_count(); }
void ofl_launch_count_reset(void) { launch_count_reset(); }
void ofl_phase_timing_enable(int on) { phase_timing_enable(on); }
int ofl_phase_timing_read(double* ms, int64_t* counts, int n, int reset) { return phase_timing_read(ms, counts, n, reset); }
int64_t ofl_perimeter_count(int64_t rows, int64_t cols) { return perimeter_count(rows, cols); }
size_t ofl_accumulation_workspace_bytes(int64_t rows, int64_t cols) { return accumulation_workspace_bytes(rows, cols); }

int ofl_flow_direction_f32(const float* dem, int64_t rows, int64_t cols, int64_t ld_dem, double nodata, uint8_t* fdr,
                           int64_t ld_fdr, int mode, int mem_kind, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0, OFL_ERR_INVALID, "negative raster size");
  OFL_REQUIRE(mode == OFL_DIR_MODE_TILE || mode == OFL_DIR_MODE_RASTER || mode == OFL_DIR_MODE_STRIP, OFL_ERR_INVALID,
              "unknown direction mode %d", mode);
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind);
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(dem != nullptr && fdr != nullptr, OFL_ERR_INVALID, "null raster pointer");
  OFL_REQUIRE(ld_dem >= cols && ld_fdr >= cols, OFL_ERR_INVALID, "leading dimension smaller than cols");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int y_off = (mode == OFL_DIR_MODE_STRIP) ? 1 : 0;
  const int64_t in_rows = rows + 2 * y_off;

  if (mem_kind == OFL_MEM_DEVICE) {
    rc = launch_direction(dem, in_rows, cols, ld_dem, nodata, fdr, rows, ld_fdr, y_off, st);
    if (rc != OFL_OK) return rc;
    if (mode == OFL_DIR_MODE_TILE) return launch_fill_border(fdr, rows, cols, ld_fdr, OFL_DIR_NODATA, st);
    return OFL_OK;
  }

  // host rasters: pitched device staging, one H2D, kernel, one D2H
  const int64_t ldd = round_up(cols, 4), ldo = round_up(cols, 16);
  void *d_dem = nullptr, *d_fdr = nullptr;
  rc = scratch_get(SCRATCH_DEM, (size_t)in_rows * ldd * sizeof(float), &d_dem);
  if (rc != OFL_OK) return rc;
  rc = scratch_get(SCRATCH_FDR, (size_t)rows * ldo, &d_fdr);
  if (rc != OFL_OK) return rc;
  if (mode != OFL_DIR_MODE_TILE)  // the ring fill of TILE mode follows the kernel: tiles take the plain path below
    return direction_from_host(dem, in_rows, rows, cols, ld_dem, nodata, y_off, static_cast<float*>(d_dem), ldd,
                               static_cast<uint8_t*>(d_fdr), ldo, fdr, ld_fdr, st, true);
  OFL_CUDA(cudaMemcpy2DAsync(d_dem, ldd * sizeof(float), dem, ld_dem * sizeof(float), cols * sizeof(float), in_rows,
                             cudaMemcpyHostToDevice, st));
  rc = launch_direction(static_cast<const float*>(d_dem), in_rows, cols, ldd, nodata, static_cast<uint8_t*>(d_fdr),
                        rows, ldo, y_off, st);
  if (rc != OFL_OK) return rc;
  if (mode == OFL_DIR_MODE_TILE) {
    rc = launch_fill_border(static_cast<uint8_t*>(d_fdr), rows, cols, ldo, OFL_DIR_NODATA, st);
    if (rc != OFL_OK) return rc;
  }
  OFL_CUDA(cudaMemcpy2DAsync(fdr, ld_fdr, d_fdr, ldo, cols, rows, cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

int ofl_flow_direction_x64(const void* dem, int elem_kind, int64_t rows, int64_t cols, int64_t ld_dem, double nodata,
                           uint8_t* fdr, int64_t ld_fdr, int mode, int mem_kind, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0, OFL_ERR_INVALID, "negative raster size");
  OFL_REQUIRE(elem_kind == OFL_ELEM_F64 || elem_kind == OFL_ELEM_I64 || elem_kind == OFL_ELEM_U64, OFL_ERR_INVALID,
              "unknown element kind %d", elem_kind);
  OFL_REQUIRE(mode == OFL_DIR_MODE_TILE || mode == OFL_DIR_MODE_RASTER || mode == OFL_DIR_MODE_STRIP, OFL_ERR_INVALID,
              "unknown direction mode %d", mode);
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind);
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(dem != nullptr && fdr != nullptr, OFL_ERR_INVALID, "null raster pointer");
  OFL_REQUIRE(ld_dem >= cols && ld_fdr >= cols, OFL_ERR_INVALID, "leading dimension smaller than cols");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int y_off = (mode == OFL_DIR_MODE_STRIP) ? 1 : 0;
  const int64_t in_rows = rows + 2 * y_off;
  const void* d_in = dem;
  uint8_t* d_out = fdr;
  int64_t ldd = ld_dem, ldo = ld_fdr;
  if (mem_kind == OFL_MEM_HOST) {
    void *b0 = nullptr, *b1 = nullptr;
    ldd = cols;
    ldo = round_up(cols, 16);
    rc = scratch_get(SCRATCH_DEM, (size_t)in_rows * ldd * 8, &b0);
    if (rc != OFL_OK) return rc;
    rc = scratch_get(SCRATCH_FDR, (size_t)rows * ldo, &b1);
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaMemcpy2DAsync(b0, ldd * 8, dem, ld_dem * 8, cols * 8, in_rows, cudaMemcpyHostToDevice, st));
    d_in = b0;
    d_out = static_cast<uint8_t*>(b1);
  }
  rc = launch_direction_generic(d_in, elem_kind, in_rows, cols, ldd, nodata, d_out, rows, ldo, y_off, st);
  if (rc != OFL_OK) return rc;
  if (mode == OFL_DIR_MODE_TILE) {
    rc = launch_fill_border(d_out, rows, cols, ldo, OFL_DIR_NODATA, st);
    if (rc != OFL_OK) return rc;
  }
  if (mem_kind == OFL_MEM_HOST) {
    OFL_CUDA(cudaMemcpy2DAsync(fdr, ld_fdr, d_out, ldo, cols, rows, cudaMemcpyDeviceToHost, st));
    OFL_CUDA(cudaStreamSynchronize(st));
  }
  return OFL_OK;
}

int ofl_fill_border_u8(uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, int value, void* stream) {
  OFL_REQUIRE(fdr != nullptr || rows * cols == 0, OFL_ERR_INVALID, "null raster pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return launch_fill_border(fdr, rows, cols, ld_fdr, value, static_cast<cudaStream_t>(stream));
}

// shared by ofl_flow_accumulation_u8 and ofl_flow_accumulation_seeded_u8
static int accumulation_entry(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, int64_t* fac, int64_t ld_fac,
                              const int64_t* perim_inflow, int64_t* perim_links, void* workspace, size_t workspace_bytes,
                              int mem_kind, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0, OFL_ERR_INVALID, "negative raster size");
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind);
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(fdr != nullptr && fac != nullptr, OFL_ERR_INVALID, "null raster pointer");
  OFL_REQUIRE(ld_fdr >= cols && ld_fac >= cols, OFL_ERR_INVALID, "leading dimension smaller than cols");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t need = accumulation_workspace_bytes(rows, cols);
  if (workspace == nullptr) {
    rc = scratch_get(SCRATCH_WORK, need, &workspace);
    if (rc != OFL_OK) return rc;
    workspace_bytes = need;
  }
  if (mem_kind == OFL_MEM_DEVICE)
    return launch_accumulation(fdr, rows, cols, ld_fdr, reinterpret_cast<long long*>(fac), ld_fac,
                               reinterpret_cast<long long*>(perim_links), workspace, workspace_bytes, st, false, false,
                               reinterpret_cast<const long long*>(perim_inflow));

  const int64_t ldi = round_up(cols, 16), ldo = round_up(cols, 2);
  void *d_fdr = nullptr, *d_fac = nullptr, *d_links = nullptr;
  rc = scratch_get(SCRATCH_FDR, (size_t)rows * ldi, &d_fdr);
  if (rc != OFL_OK) return rc;
  rc = scratch_get(SCRATCH_FAC, (size_t)rows * ldo * sizeof(int64_t), &d_fac);
  if (rc != OFL_OK) return rc;
  const int64_t n_perim = perimeter_count(rows, cols);
  if (perim_links || perim_inflow) {  // links [n][2], then the inflow [n]
    rc = scratch_get(SCRATCH_LINKS, (size_t)n_perim * 3 * sizeof(int64_t), &d_links);
    if (rc != OFL_OK) return rc;
  }
  long long* d_inflow = nullptr;
  if (perim_inflow) {
    d_inflow = static_cast<long long*>(d_links) + 2 * n_perim;
    OFL_CUDA(cudaMemcpyAsync(d_inflow, perim_inflow, (size_t)n_perim * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  }
  OFL_CUDA(cudaMemcpy2DAsync(d_fdr, ldi, fdr, ld_fdr, cols, rows, cudaMemcpyHostToDevice, st));
  rc = launch_accumulation(static_cast<const uint8_t*>(d_fdr), rows, cols, ldi, static_cast<long long*>(d_fac), ldo,
                           perim_links ? static_cast<long long*>(d_links) : nullptr, workspace, workspace_bytes, st, false,
                           false, d_inflow);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpy2DAsync(fac, ld_fac * sizeof(int64_t), d_fac, ldo * sizeof(int64_t), cols * sizeof(int64_t), rows,
                             cudaMemcpyDeviceToHost, st));
  if (perim_links)
    OFL_CUDA(cudaMemcpyAsync(perim_links, d_links, (size_t)n_perim * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

int ofl_flow_accumulation_u8(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, int64_t* fac,
                             int64_t ld_fac, int64_t* perim_links, void* workspace, size_t workspace_bytes,
                             int mem_kind, void* stream) {
  return accumulation_entry(fdr, rows, cols, ld_fdr, fac, ld_fac, nullptr, perim_links, workspace, workspace_bytes,
                            mem_kind, stream);
}

int ofl_flow_accumulation_seeded_u8(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, int64_t* fac,
                                    int64_t ld_fac, const int64_t* perim_inflow, int64_t* perim_links, void* workspace,
                                    size_t workspace_bytes, int mem_kind, void* stream) {
  return accumulation_entry(fdr, rows, cols, ld_fdr, fac, ld_fac, perim_inflow, perim_links, workspace, workspace_bytes,
                            mem_kind, stream);
}

int ofl_flow_routing_f32(const float* dem, int64_t rows, int64_t cols, int64_t ld_dem, double nodata, uint8_t* fdr,
                         int64_t ld_fdr, int64_t* fac, int64_t ld_fac, int64_t* perim_links, int mem_kind, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0, OFL_ERR_INVALID, "negative raster size");
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind);
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(dem != nullptr && fac != nullptr, OFL_ERR_INVALID, "null raster pointer");
  OFL_REQUIRE(ld_dem >= cols && ld_fac >= cols && (fdr == nullptr || ld_fdr >= cols), OFL_ERR_INVALID,
              "leading dimension smaller than cols");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* work = nullptr;
  const size_t need = accumulation_workspace_bytes(rows, cols);
  rc = scratch_get(SCRATCH_WORK, need, &work);
  if (rc != OFL_OK) return rc;
  if (mem_kind == OFL_MEM_DEVICE) {
    OFL_REQUIRE(fdr != nullptr, OFL_ERR_INVALID, "device callers provide the code raster");
    // the workspace is cleared on a second stream while the stencil runs
    rc = pipe_streams();
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaEventRecord(g_pipe.ev_in, st));
    OFL_CUDA(cudaStreamWaitEvent(g_pipe.h2d, g_pipe.ev_in, 0));
    rc = accumulation_prepare(rows, cols, work, need, g_pipe.h2d);
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaEventRecord(g_pipe.ev_prep, g_pipe.h2d));
    rc = launch_direction(dem, rows, cols, ld_dem, nodata, fdr, rows, ld_fdr, 0, st);
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaStreamWaitEvent(st, g_pipe.ev_prep, 0));
    return launch_accumulation(fdr, rows, cols, ld_fdr, reinterpret_cast<long long*>(fac), ld_fac,
                               reinterpret_cast<long long*>(perim_links), work, need, st, true, true);
  }
  const int64_t ldd = round_up(cols, 4), ldi = round_up(cols, 16), ldo = round_up(cols, 2);
  void *d_dem = nullptr, *d_fdr = nullptr, *d_fac = nullptr, *d_links = nullptr;
  rc = scratch_get(SCRATCH_DEM, (size_t)rows * ldd * sizeof(float), &d_dem);
  if (rc != OFL_OK) return rc;
  rc = scratch_get(SCRATCH_FDR, (size_t)rows * ldi, &d_fdr);
  if (rc != OFL_OK) return rc;
  rc = scratch_get(SCRATCH_FAC, (size_t)rows * ldo * sizeof(int64_t), &d_fac);
  if (rc != OFL_OK) return rc;
  const int64_t n_perim = perimeter_count(rows, cols);
  if (perim_links) {
    rc = scratch_get(SCRATCH_LINKS, (size_t)n_perim * 2 * sizeof(int64_t), &d_links);
    if (rc != OFL_OK) return rc;
  }
  // codes stream back to the host while later bands are still uploading; the counts follow the accumulation
  rc = direction_from_host(dem, rows, rows, cols, ld_dem, nodata, 0, static_cast<float*>(d_dem), ldd,
                           static_cast<uint8_t*>(d_fdr), ldi, fdr, ld_fdr, st, false);
  if (rc != OFL_OK) return rc;
  rc = launch_accumulation(static_cast<const uint8_t*>(d_fdr), rows, cols, ldi, static_cast<long long*>(d_fac), ldo,
                           static_cast<long long*>(d_links), work, need, st, false, true);
  if (rc == OFL_OK) {
    cudaError_t e = cudaMemcpy2DAsync(fac, ld_fac * sizeof(int64_t), d_fac, ldo * sizeof(int64_t), cols * sizeof(int64_t),
                                      rows, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && perim_links)
      e = cudaMemcpyAsync(perim_links, d_links, (size_t)n_perim * 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpyAsync(fac)", __FILE__, __LINE__);
  }
  cudaStreamSynchronize(st);
  cudaStreamSynchronize(g_pipe.d2h);
  return rc;
}

int ofl_check_accumulation_u8(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld_fdr, const int64_t* fac,
                              int64_t ld_fac, int64_t* n_bad, int mem_kind, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0 && n_bad != nullptr, OFL_ERR_INVALID, "bad argument");
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind);
  *n_bad = 0;
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(fdr != nullptr && fac != nullptr, OFL_ERR_INVALID, "null raster pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* d_cnt = nullptr;
  rc = scratch_get(SCRATCH_MISC, 256, &d_cnt);
  if (rc != OFL_OK) return rc;
  const uint8_t* dfdr = fdr;
  const long long* dfac = reinterpret_cast<const long long*>(fac);
  int64_t ldi = ld_fdr, ldo = ld_fac;
  if (mem_kind == OFL_MEM_HOST) {
    void *b0 = nullptr, *b1 = nullptr;
    ldi = round_up(cols, 16);
    ldo = cols;
    rc = scratch_get(SCRATCH_FDR, (size_t)rows * ldi, &b0);
    if (rc != OFL_OK) return rc;
    rc = scratch_get(SCRATCH_FAC, (size_t)rows * ldo * sizeof(int64_t), &b1);
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaMemcpy2DAsync(b0, ldi, fdr, ld_fdr, cols, rows, cudaMemcpyHostToDevice, st));
    OFL_CUDA(cudaMemcpy2DAsync(b1, ldo * sizeof(int64_t), fac, ld_fac * sizeof(int64_t), cols * sizeof(int64_t), rows,
                               cudaMemcpyHostToDevice, st));
    dfdr = static_cast<const uint8_t*>(b0);
    dfac = static_cast<const long long*>(b1);
  }
  rc = launch_check(dfdr, rows, cols, ldi, dfac, ldo, static_cast<unsigned long long*>(d_cnt), st);
  if (rc != OFL_OK) return rc;
  unsigned long long h = 0;
  OFL_CUDA(cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  *n_bad = (int64_t)h;
  return OFL_OK;
}

int ofl_strip_check_accumulation_u8(const uint8_t* fdr_halo, int64_t rows, int64_t cols, int64_t ld_fdr,
                                    const int64_t* fac, int64_t ld_fac, const int64_t* fac_above,
                                    const int64_t* fac_below, int64_t* n_bad, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0 && n_bad != nullptr, OFL_ERR_INVALID, "bad argument");
  *n_bad = 0;
  if (rows == 0 || cols == 0) return OFL_OK;
  OFL_REQUIRE(fdr_halo != nullptr && fac != nullptr, OFL_ERR_INVALID, "null raster pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* d_cnt = nullptr;
  rc = scratch_get(SCRATCH_MISC, 256, &d_cnt);
  if (rc != OFL_OK) return rc;
  rc = launch_check(fdr_halo, rows, cols, ld_fdr, reinterpret_cast<const long long*>(fac), ld_fac,
                    static_cast<unsigned long long*>(d_cnt), st, 1, reinterpret_cast<const long long*>(fac_above),
                    reinterpret_cast<const long long*>(fac_below));
  if (rc != OFL_OK) return rc;
  unsigned long long h = 0;
  OFL_CUDA(cudaMemcpyAsync(&h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  *n_bad = (int64_t)h;
  return OFL_OK;
}

// ---------------------------------------------------------------- flat resolution (csrc/flats.cu)
size_t ofl_flats_workspace_bytes(int64_t rows, int64_t cols) {
  return (rows > 0 && cols > 0) ? flats_workspace_bytes(rows, cols) : 0;
}

#define OFL_FLATS_COMMON(...)                                                                                        \
  OFL_REQUIRE(rows >= 0 && cols >= 0, OFL_ERR_INVALID, "negative raster size");                                       \
  OFL_REQUIRE(mem_kind == OFL_MEM_HOST || mem_kind == OFL_MEM_DEVICE, OFL_ERR_INVALID, "unknown mem_kind %d", mem_kind); \
  if (rows == 0 || cols == 0) return OFL_OK;                                                                           \
  OFL_REQUIRE(__VA_ARGS__, OFL_ERR_INVALID, "null raster pointer");                                                    \
  int rc = ensure_init();                                                                                              \
  if (rc != OFL_OK) return rc;                                                                                         \
  cudaStream_t st = static_cast<cudaStream_t>(stream);                                                                 \
  const size_t n = (size_t)rows * (size_t)cols

int ofl_flat_edges_f32(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, uint8_t* edges, int64_t* n_low,
                       int64_t* n_high, int mem_kind, void* stream) {
  if (n_low) *n_low = 0;
  if (n_high) *n_high = 0;
  OFL_FLATS_COMMON(dem && fdr && edges);
  void* d_cnt = nullptr;
  rc = scratch_get(SCRATCH_MISC, 256, &d_cnt);
  if (rc != OFL_OK) return rc;
  if (mem_kind == OFL_MEM_DEVICE)
    return launch_flat_edges(dem, fdr, rows, cols, edges, n_low, n_high, static_cast<unsigned*>(d_cnt), st);
  void *d_dem = nullptr, *d_fdr = nullptr, *d_edges = nullptr;
  if ((rc = scratch_get(SCRATCH_DEM, n * sizeof(float), &d_dem)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FDR, n, &d_fdr)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_LINKS, n, &d_edges)) != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(d_dem, dem, n * sizeof(float), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_fdr, fdr, n, cudaMemcpyHostToDevice, st));
  rc = launch_flat_edges(static_cast<float*>(d_dem), static_cast<uint8_t*>(d_fdr), rows, cols,
                         static_cast<uint8_t*>(d_edges), n_low, n_high, static_cast<unsigned*>(d_cnt), st);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(edges, d_edges, n, cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

// resolve (+ optionally rewrite the codes): shared by ofl_resolve_flats_f32 and ofl_fix_flats_f32
static int flats_run(const float* dem, const uint8_t* fdr_in, uint8_t* fdr_out, int64_t rows, int64_t cols,
                     int32_t* flat_mask, int32_t* labels, int64_t* info, void* workspace, size_t workspace_bytes,
                     int mem_kind, cudaStream_t st) {
  const size_t n = (size_t)rows * (size_t)cols;
  int rc;
  if (!workspace) {
    workspace_bytes = flats_workspace_bytes(rows, cols);
    if ((rc = scratch_get(SCRATCH_WORK, workspace_bytes, &workspace)) != OFL_OK) return rc;
  }
  if (mem_kind == OFL_MEM_DEVICE) {
    OFL_REQUIRE(flat_mask && labels, OFL_ERR_INVALID, "device callers provide flat_mask and labels");
    rc = launch_resolve_flats(dem, fdr_in, rows, cols, flat_mask, labels, info, workspace, workspace_bytes, st);
    if (rc == OFL_OK && fdr_out) rc = launch_masked_flow_dirs(flat_mask, labels, fdr_out, rows, cols, st);
    return rc;
  }
  void *d_dem = nullptr, *d_fdr = nullptr, *d_out = nullptr;
  if ((rc = scratch_get(SCRATCH_DEM, n * sizeof(float), &d_dem)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FDR, n, &d_fdr)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FAC, 2 * n * sizeof(int32_t), &d_out)) != OFL_OK) return rc;
  int32_t* d_mask = static_cast<int32_t*>(d_out);
  int32_t* d_labels = d_mask + n;
  OFL_CUDA(cudaMemcpyAsync(d_dem, dem, n * sizeof(float), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_fdr, fdr_in, n, cudaMemcpyHostToDevice, st));
  rc = launch_resolve_flats(static_cast<float*>(d_dem), static_cast<uint8_t*>(d_fdr), rows, cols, d_mask, d_labels, info,
                            workspace, workspace_bytes, st);
  if (rc != OFL_OK) return rc;
  if (fdr_out) {
    rc = launch_masked_flow_dirs(d_mask, d_labels, static_cast<uint8_t*>(d_fdr), rows, cols, st);
    if (rc != OFL_OK) return rc;
    OFL_CUDA(cudaMemcpyAsync(fdr_out, d_fdr, n, cudaMemcpyDeviceToHost, st));
  }
  if (flat_mask) OFL_CUDA(cudaMemcpyAsync(flat_mask, d_mask, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (labels) OFL_CUDA(cudaMemcpyAsync(labels, d_labels, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

int ofl_resolve_flats_f32(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, int32_t* flat_mask,
                          int32_t* labels, int64_t* info, void* workspace, size_t workspace_bytes, int mem_kind,
                          void* stream) {
  if (info) for (int k = 0; k < 5; ++k) info[k] = 0;
  OFL_FLATS_COMMON(dem && fdr && flat_mask && labels);
  (void)n;
  return flats_run(dem, fdr, nullptr, rows, cols, flat_mask, labels, info, workspace, workspace_bytes, mem_kind, st);
}

int ofl_fix_flats_f32(const float* dem, uint8_t* fdr, int64_t rows, int64_t cols, int32_t* flat_mask, int32_t* labels,
                      int64_t* info, void* workspace, size_t workspace_bytes, int mem_kind, void* stream) {
  if (info) for (int k = 0; k < 5; ++k) info[k] = 0;
  OFL_FLATS_COMMON(dem && fdr);
  (void)n;
  return flats_run(dem, fdr, fdr, rows, cols, flat_mask, labels, info, workspace, workspace_bytes, mem_kind, st);
}

int ofl_d8_masked_flow_dirs_i32(const int32_t* flat_mask, const int32_t* labels, uint8_t* fdr, int64_t rows, int64_t cols,
                                int mem_kind, void* stream) {
  OFL_FLATS_COMMON(flat_mask && labels && fdr);
  if (mem_kind == OFL_MEM_DEVICE) return launch_masked_flow_dirs(flat_mask, labels, fdr, rows, cols, st);
  void *d_fdr = nullptr, *d_in = nullptr;
  if ((rc = scratch_get(SCRATCH_FDR, n, &d_fdr)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FAC, 2 * n * sizeof(int32_t), &d_in)) != OFL_OK) return rc;
  int32_t* d_mask = static_cast<int32_t*>(d_in);
  int32_t* d_labels = d_mask + n;
  OFL_CUDA(cudaMemcpyAsync(d_mask, flat_mask, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_labels, labels, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_fdr, fdr, n, cudaMemcpyHostToDevice, st));
  rc = launch_masked_flow_dirs(d_mask, d_labels, static_cast<uint8_t*>(d_fdr), rows, cols, st);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(fdr, d_fdr, n, cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

int ofl_flat_gradient_i32(const int32_t* labels, const uint8_t* fdr, int64_t rows, int64_t cols, const int32_t* seeds,
                          int64_t n_seeds, int towards, int32_t* flat_mask, int32_t* flat_height, int64_t n_heights,
                          void* workspace, size_t workspace_bytes, int mem_kind, void* stream) {
  OFL_FLATS_COMMON(labels && fdr && flat_mask && (seeds || n_seeds == 0) && (flat_height || n_heights == 0));
  OFL_REQUIRE(n_seeds >= 0 && n_heights >= 0 && (size_t)n_seeds <= n && (size_t)n_heights <= n, OFL_ERR_INVALID,
              "seed / flat_height count out of range");
  if (!workspace) {
    workspace_bytes = flats_workspace_bytes(rows, cols);
    if ((rc = scratch_get(SCRATCH_WORK, workspace_bytes, &workspace)) != OFL_OK) return rc;
  }
  if (mem_kind == OFL_MEM_DEVICE)
    return launch_flat_gradient(labels, fdr, rows, cols, seeds, n_seeds, towards, flat_mask, flat_height, n_heights,
                                workspace, workspace_bytes, st);
  void *d_fdr = nullptr, *d_io = nullptr, *d_small = nullptr;
  if ((rc = scratch_get(SCRATCH_FDR, n, &d_fdr)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FAC, 2 * n * sizeof(int32_t), &d_io)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_LINKS, (size_t)(n_seeds + n_heights + 2) * sizeof(int32_t), &d_small)) != OFL_OK) return rc;
  int32_t* d_mask = static_cast<int32_t*>(d_io);
  int32_t* d_labels = d_mask + n;
  int32_t* d_seeds = static_cast<int32_t*>(d_small);
  int32_t* d_fh = d_seeds + n_seeds;
  OFL_CUDA(cudaMemcpyAsync(d_mask, flat_mask, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_labels, labels, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  OFL_CUDA(cudaMemcpyAsync(d_fdr, fdr, n, cudaMemcpyHostToDevice, st));
  if (n_seeds) OFL_CUDA(cudaMemcpyAsync(d_seeds, seeds, (size_t)n_seeds * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (n_heights) OFL_CUDA(cudaMemcpyAsync(d_fh, flat_height, (size_t)n_heights * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  rc = launch_flat_gradient(d_labels, static_cast<uint8_t*>(d_fdr), rows, cols, d_seeds, n_seeds, towards, d_mask, d_fh,
                            n_heights, workspace, workspace_bytes, st);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpyAsync(flat_mask, d_mask, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (n_heights) OFL_CUDA(cudaMemcpyAsync(flat_height, d_fh, (size_t)n_heights * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

// ---------------------------------------------------------------- single-cell pit breaching (csrc/pits.cu)
size_t ofl_pits_workspace_bytes(int64_t rows, int64_t cols) {
  return (rows > 0 && cols > 0) ? pits_workspace_bytes(rows, cols) : 0;
}

int ofl_breach_single_cell_pits_f32(float* chunk, int64_t rows, int64_t cols, int64_t ld_chunk, double nodata,
                                    int8_t* unsolved, int64_t* info, void* workspace, size_t workspace_bytes, int mem_kind,
                                    void* stream) {
  if (info) info[0] = info[1] = info[2] = 0;
  OFL_FLATS_COMMON(chunk && unsolved);
  OFL_REQUIRE(ld_chunk >= cols, OFL_ERR_INVALID, "leading dimension smaller than cols");
  if (!workspace) {
    workspace_bytes = pits_workspace_bytes(rows, cols);
    if ((rc = scratch_get(SCRATCH_WORK, workspace_bytes, &workspace)) != OFL_OK) return rc;
  }
  if (mem_kind == OFL_MEM_DEVICE)
    return launch_breach_pits(chunk, rows, cols, ld_chunk, nodata, unsolved, info, workspace, workspace_bytes, st);
  void *d_chunk = nullptr, *d_uns = nullptr;
  if ((rc = scratch_get(SCRATCH_DEM, n * sizeof(float), &d_chunk)) != OFL_OK) return rc;
  if ((rc = scratch_get(SCRATCH_FDR, n, &d_uns)) != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpy2DAsync(d_chunk, cols * sizeof(float), chunk, ld_chunk * sizeof(float), cols * sizeof(float), rows,
                             cudaMemcpyHostToDevice, st));
  rc = launch_breach_pits(static_cast<float*>(d_chunk), rows, cols, cols, nodata, static_cast<int8_t*>(d_uns), info,
                          workspace, workspace_bytes, st);
  if (rc != OFL_OK) return rc;
  OFL_CUDA(cudaMemcpy2DAsync(chunk, ld_chunk * sizeof(float), d_chunk, cols * sizeof(float), cols * sizeof(float), rows,
                             cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaMemcpyAsync(unsolved, d_uns, n, cudaMemcpyDeviceToHost, st));
  OFL_CUDA(cudaStreamSynchronize(st));
  return OFL_OK;
}

size_t ofl_strip_workspace_bytes(int64_t rows, int64_t cols) { return strip_workspace_bytes(rows, cols); }
size_t ofl_strip_boundary_workspace_bytes(int n_strips, int64_t cols) {
  return strip_boundary_workspace_bytes(n_strips, cols);
}

size_t ofl_strip_record_bytes(int64_t cols) { return strip_record_bytes(cols); }

int ofl_strip_accum_local(const uint8_t* fdr_halo, int64_t rows, int64_t cols, int64_t ld_fdr, int has_above,
                          int has_below, int64_t* fac, int64_t ld_fac, void* workspace, size_t workspace_bytes,
                          void* record, void* stream) {
  OFL_REQUIRE(fdr_halo && fac && workspace && record, OFL_ERR_INVALID, "null pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return strip_accum_local(fdr_halo, rows, cols, ld_fdr, has_above, has_below, reinterpret_cast<long long*>(fac), ld_fac,
                           workspace, workspace_bytes, record, static_cast<cudaStream_t>(stream));
}

int ofl_strip_boundary_solve(const void* records_all, int n_strips, int64_t cols, int64_t* J_all, void* workspace,
                             size_t workspace_bytes, void* stream) {
  OFL_REQUIRE(records_all && J_all && workspace, OFL_ERR_INVALID, "null pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return strip_boundary_solve(records_all, n_strips, cols, reinterpret_cast<long long*>(J_all), workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
}

int ofl_strip_collect_flags(const void* strip_workspace, int64_t rows, int64_t cols, const void* boundary_workspace,
                            int n_strips, int32_t* flags, void* stream) {
  OFL_REQUIRE(flags != nullptr, OFL_ERR_INVALID, "null pointer");
  OFL_REQUIRE(strip_workspace == nullptr || (rows >= 1 && cols >= 1), OFL_ERR_INVALID, "bad strip size");
  OFL_REQUIRE(boundary_workspace == nullptr || (n_strips >= 1 && cols >= 1), OFL_ERR_INVALID, "bad boundary graph size");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return strip_collect_flags(strip_workspace, rows, cols, boundary_workspace, n_strips, flags,
                             static_cast<cudaStream_t>(stream));
}

int ofl_strip_accum_final(const uint8_t* fdr_halo, int64_t rows, int64_t cols, int64_t ld_fdr, int has_above,
                          int has_below, const int64_t* J_mine, void* workspace, size_t workspace_bytes, int64_t* fac,
                          int64_t ld_fac, void* stream) {
  OFL_REQUIRE(fdr_halo && fac && workspace && J_mine, OFL_ERR_INVALID, "null pointer");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return strip_accum_final(fdr_halo, rows, cols, ld_fdr, has_above, has_below, reinterpret_cast<const long long*>(J_mine),
                           workspace, workspace_bytes, reinterpret_cast<long long*>(fac), ld_fac,
                           static_cast<cudaStream_t>(stream));
}

int ofl_synth_dem_f32(float* dem, int64_t rows, int64_t cols, int64_t ld_dem, int64_t row0, int64_t total_rows,
                      uint64_t seed, int kind, float relief, int holes_permille, float nodata, void* stream) {
  OFL_REQUIRE(rows >= 0 && cols >= 0 && ld_dem >= cols, OFL_ERR_INVALID, "bad raster size");
  OFL_REQUIRE(dem != nullptr || rows * cols == 0, OFL_ERR_INVALID, "null raster pointer");
  OFL_REQUIRE(kind >= 0 && kind <= 4, OFL_ERR_INVALID, "unknown synthetic DEM kind %d", kind);
  OFL_REQUIRE((kind != 3 && kind != 4) || total_rows / 2 * cols < 0xFD000000ll, OFL_ERR_INVALID,
              "serpentine DEM: the chain would run out of finite float32 values");
  int rc = ensure_init();
  if (rc != OFL_OK) return rc;
  return launch_synth(dem, rows, cols, ld_dem, row0, total_rows, seed, kind, relief, holes_permille, nodata,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
