/*
 * synth_host.c -- host restatement of the benchmark DEM generator (overflow_b200/csrc/synth.cu).
 *
 * TEST INFRASTRUCTURE ONLY (see d8_oracle.c).  bench.py's reference arm and the parity windows need
 * the synthetic benchmark DEM without mapping the CUDA library into the process.  This is not a
 * restatement of anything in the reference (the reference ships no generator); it follows synth.cu
 * operation for operation in float32 / uint32 (built with -ffp-contract=off, the device side with
 * --fmad=false), so the arrays are bit-identical to the device's (tests/test_gpu_synth.py) and to the
 * numpy form in oracle/synth.py (tests/test_synth_host.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t sh_hash3(uint32_t x, uint32_t y, uint32_t s) {
  uint32_t h = x * 0x9E3779B1u ^ (y * 0x85EBCA77u + 0x165667B1u) ^ (s * 0xC2B2AE3Du);
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}

static inline float sh_lattice(int x, int y, uint32_t s) {
  return (float)(sh_hash3((uint32_t)x, (uint32_t)y, s) >> 8) * (1.0f / 16777216.0f);
}

static float sh_vnoise(float fx, float fy, uint32_t s) {
  const float flx = floorf(fx), fly = floorf(fy);
  const int ix = (int)flx, iy = (int)fly;
  float tx = fx - flx, ty = fy - fly;
  tx = tx * tx * (3.f - 2.f * tx);
  ty = ty * ty * (3.f - 2.f * ty);
  const float a = sh_lattice(ix, iy, s), b = sh_lattice(ix + 1, iy, s);
  const float c = sh_lattice(ix, iy + 1, s), d = sh_lattice(ix + 1, iy + 1, s);
  const float top = a + (b - a) * tx, bot = c + (d - c) * tx;
  return top + (bot - top) * ty;
}

static float sh_serpentine(int64_t gr, int64_t c, int64_t total_rows, int64_t cols) {
  const float wall = 8.0e37f;
  if (gr <= 0 || gr >= total_rows - 1 || c <= 0 || c >= cols - 1) return wall;
  const int64_t wc = cols - 2, n_runs = (total_rows - 1) / 2;
  int64_t k;
  if (gr & 1) {
    const int64_t m = (gr - 1) >> 1;
    k = m * (wc + 1) + ((m & 1) ? (cols - 2 - c) : (c - 1));
  } else {
    const int64_t m = (gr >> 1) - 1;
    if (m + 1 >= n_runs) return wall;
    if (c != ((m & 1) ? 1 : cols - 2)) return wall;
    k = m * (wc + 1) + wc;
  }
  const long long ord = 0x7E000000ll - (long long)k;
  const uint32_t bits = ord >= 0 ? (uint32_t)ord : (0x80000000u | (uint32_t)(-ord));
  float z;
  memcpy(&z, &bits, 4);
  return z;
}

/* rows row0 .. row0+rows, columns col0 .. col0+cols of the total_rows x total_cols raster */
void orc_synth_dem_f32(float* dem, int64_t rows, int64_t cols, int64_t ld, int64_t row0, int64_t col0,
                       int64_t total_rows, int64_t total_cols, uint64_t seed64, int kind, float relief,
                       int holes_permille, float nodata) {
  const uint32_t seed = (uint32_t)(seed64 ^ (seed64 >> 32));
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const int64_t gr = row0 + r;
    for (int64_t cc = 0; cc < cols; ++cc) {
      const int64_t c = col0 + cc;
      float z;
      if (gr < 0 || gr >= total_rows) {
        z = nodata;
      } else if (kind == 2) {
        z = (float)(total_rows - 1 - gr) + 0.25f * (float)(total_cols - 1 - c);
      } else if (kind == 3) {
        z = sh_serpentine(gr, c, total_rows, total_cols);
      } else if (kind == 4) {
        z = sh_serpentine(c, gr, total_cols, total_rows);
      } else {
        float amp = 1.f, sum = 0.f, norm = 0.f, freq = 1.0f / 4096.0f;
        for (int o = 0; o < 12; ++o) {
          sum += amp * sh_vnoise((float)c * freq, (float)gr * freq, seed + 31u * (uint32_t)o);
          norm += amp;
          amp *= 0.55f;
          freq *= 2.f;
        }
        z = relief * sum / norm;
        if (kind == 1) z = floorf(z);
        if (holes_permille > 0) {
          const uint32_t hb = sh_hash3((uint32_t)(c >> 8), (uint32_t)(gr >> 8), seed ^ 0xA5A5A5A5u);
          if ((int)(hb % 1000u) < holes_permille * 16) {
            const int hx = (int)((hb >> 10) & 127), hy = (int)((hb >> 17) & 127);
            const int cx = (int)(c & 255), cy = (int)(gr & 255);
            if (cx >= hx && cx < hx + 64 && cy >= hy && cy < hy + 64) z = nodata;
          }
        }
      }
      dem[r * ld + cc] = z;
    }
  }
}
