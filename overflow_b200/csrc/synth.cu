// synth.cu -- seeded synthetic DEMs generated straight into device memory (bench / large parity runs).
// Every cell is a pure function of (global row, column, seed): row strips on different GPUs
// synthesise identical terrain for their rows and halos.
#include "common.cuh"

namespace ofl {

__device__ __forceinline__ uint32_t hash3(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0x165667B1u) ^ (s * 0xC2B2AE3Du);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}

__device__ __forceinline__ float lattice(int x, int y, uint32_t s) {
  return (float)(hash3((uint32_t)x, (uint32_t)y, s) >> 8) * (1.0f / 16777216.0f);
}

// value noise with smoothstep interpolation
__device__ float vnoise(float fx, float fy, uint32_t s) {
  const float flx = floorf(fx), fly = floorf(fy);
  const int ix = (int)flx, iy = (int)fly;
  float tx = fx - flx, ty = fy - fly;
  tx = tx * tx * (3.f - 2.f * tx);
  ty = ty * ty * (3.f - 2.f * ty);
  const float a = lattice(ix, iy, s), b = lattice(ix + 1, iy, s);
  const float c = lattice(ix, iy + 1, s), d = lattice(ix + 1, iy + 1, s);
  const float top = a + (b - a) * tx, bot = c + (d - c) * tx;
  return top + (bot - top) * ty;
}

// Walled serpentine (SURVEY 8d config 5 ii): ONE channel of about rows*cols/2 cells.  Channel rows are the odd
// rows 1, 3, ..; it runs east on rows 1, 5, .. and west on rows 3, 7, .. over columns 1 .. cols-2 and steps down
// through a one-cell gap in the wall row at the end of each run.  Everything else is wall (8e37), which drains
// into the channel; the raster's outer ring is wall, so the channel ends in an interior pit.  The k-th channel
// cell holds the k-th float32 below 0x7E000000 in the total order of finite floats (bit patterns walked down
// through +0 into the negative denormals, -0 skipped): strictly decreasing along the chain for up to 4.2e9
// cells, every difference between consecutive cells exact (one ulp), no two cells equal.
constexpr float SERP_WALL = 8.0e37f;  // four of them still add up to a finite float32
constexpr long long SERP_ORD0 = 0x7E000000ll;  // 4.25e37, below the walls

__device__ __forceinline__ float serpentine_cell(int64_t gr, int64_t c, int64_t total_rows, int64_t cols) {
  if (gr <= 0 || gr >= total_rows - 1 || c <= 0 || c >= cols - 1) return SERP_WALL;
  const int64_t wc = cols - 2;                    // channel cells per run
  const int64_t n_runs = (total_rows - 1) / 2;    // channel rows 1, 3, .., <= total_rows - 2
  int64_t k;
  if (gr & 1) {
    const int64_t m = (gr - 1) >> 1;
    const int64_t pos = (m & 1) ? (cols - 2 - c) : (c - 1);
    k = m * (wc + 1) + pos;
  } else {
    const int64_t m = (gr >> 1) - 1;              // the run above this wall row
    if (m + 1 >= n_runs) return SERP_WALL;        // no run below: no gap
    const int64_t gap_col = (m & 1) ? 1 : cols - 2;
    if (c != gap_col) return SERP_WALL;
    k = m * (wc + 1) + wc;
  }
  const long long ord = SERP_ORD0 - (long long)k;
  const uint32_t bits = ord >= 0 ? (uint32_t)ord : (0x80000000u | (uint32_t)(-ord));
  return __uint_as_float(bits);
}

__global__ void synth_kernel(float* dem, int64_t rows, int64_t cols, int64_t ld, int64_t row0, int64_t total_rows,
                             uint32_t seed, int kind, float relief, int holes_permille, float nodata) {
  const int64_t n = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const int64_t gr = row0 + r;
    float z;
    if (gr < 0 || gr >= total_rows) {
      z = nodata;  // halo row outside the raster
    } else if (kind == 2) {
      // tilted plane: drains towards the south-east corner, all values exact in float32
      z = (float)(total_rows - 1 - gr) + 0.25f * (float)(cols - 1 - c);
    } else if (kind == 3) {
      z = serpentine_cell(gr, c, total_rows, cols);
    } else if (kind == 4) {
      z = serpentine_cell(c, gr, cols, total_rows);  // the same channel transposed: runs north-south, crossing every
                                                     // row-strip boundary cols / 2 times
    } else {
      // 12 octaves, persistence 0.55: rough fractal relief (many local pits, like beta ~ 2 spectra)
      float amp = 1.f, sum = 0.f, norm = 0.f, freq = 1.0f / 4096.0f;
      for (int o = 0; o < 12; ++o) {
        sum += amp * vnoise((float)c * freq, (float)gr * freq, seed + 31u * (uint32_t)o);
        norm += amp;
        amp *= 0.55f;
        freq *= 2.f;
      }
      z = relief * sum / norm;
      if (kind == 1) z = floorf(z);  // terraces: large flats
      if (holes_permille > 0) {
        // nodata rectangles on a 256-cell lattice: a block is a hole with probability holes_permille/1000 * 16
        const uint32_t hb = hash3((uint32_t)(c >> 8), (uint32_t)(gr >> 8), seed ^ 0xA5A5A5A5u);
        if ((int)(hb % 1000u) < holes_permille * 16) {
          const int hx = (int)((hb >> 10) & 127), hy = (int)((hb >> 17) & 127);
          const int cx = (int)(c & 255), cy = (int)(gr & 255);
          if (cx >= hx && cx < hx + 64 && cy >= hy && cy < hy + 64) z = nodata;
        }
      }
    }
    dem[r * ld + c] = z;
  }
}

int launch_synth(float* dem, int64_t rows, int64_t cols, int64_t ld, int64_t row0, int64_t total_rows, uint64_t seed,
                 int kind, float relief, int holes_permille, float nodata, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return OFL_OK;
  const int64_t n = rows * cols;
  const int64_t want = (n + 255) / 256;
  const int blocks = (int)(want < (int64_t)sm_count() * 32 ? want : (int64_t)sm_count() * 32);
  synth_kernel<<<blocks, 256, 0, st>>>(dem, rows, cols, ld, row0, total_rows, (uint32_t)(seed ^ (seed >> 32)), kind,
                                       relief, holes_permille, nodata);
  OFL_CHECK_LAUNCH();
  return OFL_OK;
}

}  // namespace ofl
