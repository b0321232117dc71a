/*
 * d8_oracle.c -- CPU restatement of overflow's D8 flow-routing hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path in
 * overflow_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, load or call it.  Nothing under
 * overflow_b200/ links or imports it and the product path never falls back to it.
 *
 * Parity status: PINNED.  oracle/gen_golden.py imports the reference's own numba
 * kernels from /root/reference and records their outputs under tests/golden/;
 * tests/test_oracle.py checks every function below against those fixtures
 * (including the reference's two known-answer tests).
 *
 * Each function cites the reference lines it restates (paths relative to
 * /root/reference/).  The restatement is written from the algorithm's behaviour,
 * in C, with explicit types where the reference relies on numba's typing:
 *   - elevation differences are taken in float32, then widened and divided in
 *     float64 (src/overflow/flow_direction.py:94-96; numba types binop_sub as
 *     float32 and binop_truediv as float64),
 *   - NEIGHBOR_OFFSETS[code] for code >= 8 is an out-of-bounds read in the
 *     reference that lands "outside the tile" (src/overflow/flow_accumulation.py:27-37);
 *     here that is stated explicitly as "codes >= 8 have no downstream cell".
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off, no fast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_DIR_UNDEFINED 8 /* src/overflow/constants.py:23 */
#define ORC_DIR_NODATA 9    /* src/overflow/constants.py:24 */
#define ORC_FAC_NODATA (-9999) /* src/overflow/constants.py:57 */

/* scan order E, NE, N, NW, W, SW, S, SE -- src/overflow/constants.py:29-40 */
static const int ORC_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
static const int ORC_DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------------------
 * calculate_slope -- src/overflow/flow_direction.py:72-96
 * neighbour == nodata (compared after widening to double) -> +inf; otherwise the
 * float32 difference widened to double and divided by sqrt(2) or 1.
 * ------------------------------------------------------------------------- */
static inline double orc_slope_f32(float z, float n, int diagonal, double nodata) {
  if ((double)n == nodata) return INFINITY;
  volatile float diff = z - n; /* float32 subtraction, rounded once */
  return (double)diff / (diagonal ? sqrt(2.0) : 1.0);
}

static inline double orc_slope_f64(double z, double n, int diagonal, double nodata) {
  if (n == nodata) return INFINITY;
  return (z - n) / (diagonal ? sqrt(2.0) : 1.0);
}

/* ---------------------------------------------------------------------------
 * flow_direction_for_tile -- src/overflow/flow_direction.py:14-69
 * Interior cells only (rows 1..R-2, cols 1..C-2).  The reference leaves the
 * border ring uninitialised (np.empty, :39); here it is set to `border`.
 * prange over rows (:47) -> omp parallel for.
 * ------------------------------------------------------------------------- */
#define ORC_DEFINE_DIRECTION(NAME, T, SLOPE)                                          \
  void NAME(const T* dem, int64_t rows, int64_t cols, int64_t ld, double nodata,      \
            uint8_t* fdr, int64_t ld_out, int border) {                               \
    for (int64_t r = 0; r < rows; ++r) {                                              \
      if (r == 0 || r == rows - 1) {                                                  \
        for (int64_t c = 0; c < cols; ++c) fdr[r * ld_out + c] = (uint8_t)border;     \
      } else if (cols > 0) {                                                          \
        fdr[r * ld_out] = (uint8_t)border;                                            \
        fdr[r * ld_out + cols - 1] = (uint8_t)border;                                 \
      }                                                                               \
    }                                                                                 \
    _Pragma("omp parallel for schedule(static)")                                      \
    for (int64_t r = 1; r < rows - 1; ++r) {                                          \
      for (int64_t c = 1; c < cols - 1; ++c) {                                        \
        T z = dem[r * ld + c];                                                        \
        if ((double)z != nodata) {                                                    \
          double max_slope = -INFINITY;                                               \
          int max_index = -1;                                                         \
          int all_non_positive = 1;                                                   \
          for (int i = 0; i < 8; ++i) {                                               \
            T n = dem[(r + ORC_DY[i]) * ld + (c + ORC_DX[i])];                        \
            double s = SLOPE(z, n, (ORC_DY[i] != 0 && ORC_DX[i] != 0), nodata);       \
            if (s > max_slope) {                                                      \
              max_slope = s;                                                          \
              max_index = i;                                                          \
            }                                                                         \
            if (s > 0) all_non_positive = 0;                                          \
          }                                                                           \
          fdr[r * ld_out + c] =                                                       \
              all_non_positive ? ORC_DIR_UNDEFINED : (uint8_t)max_index;              \
        } else {                                                                      \
          fdr[r * ld_out + c] = ORC_DIR_NODATA;                                       \
        }                                                                             \
      }                                                                               \
    }                                                                                 \
  }

/* Integer DEMs: numba takes the difference in int64 for signed element types and in uint64 for
 * unsigned ones -- an uphill neighbour wraps around to ~2^64 -- and widens it to float64 for the
 * division (flow_direction.py:94-96; typing probed with numba 0.65).  The nodata test compares the
 * element widened to float64 (:91). */
static inline double orc_slope_i64(int64_t z, int64_t n, int diagonal, double nodata) {
  if ((double)n == nodata) return INFINITY;
  int64_t diff = (int64_t)((uint64_t)z - (uint64_t)n); /* two's-complement wrap, like int64 */
  return (double)diff / (diagonal ? sqrt(2.0) : 1.0);
}

static inline double orc_slope_u64(uint64_t z, uint64_t n, int diagonal, double nodata) {
  if ((double)n == nodata) return INFINITY;
  uint64_t diff = z - n; /* modulo 2^64 */
  return (double)diff / (diagonal ? sqrt(2.0) : 1.0);
}

ORC_DEFINE_DIRECTION(orc_flow_direction_f32, float, orc_slope_f32)
ORC_DEFINE_DIRECTION(orc_flow_direction_f64, double, orc_slope_f64)
ORC_DEFINE_DIRECTION(orc_flow_direction_i64, int64_t, orc_slope_i64)
ORC_DEFINE_DIRECTION(orc_flow_direction_u64, uint64_t, orc_slope_u64)

/* ---------------------------------------------------------------------------
 * get_next_cell -- src/overflow/flow_accumulation.py:13-37
 * Returns 1 and (*nr,*nc,*nv) when the downstream cell lies inside the tile.
 * Returns 0 ("outside": value NODATA in the reference) when it does not, which
 * includes every code >= 8 (the reference's out-of-bounds offset read).
 * ------------------------------------------------------------------------- */
static inline int orc_next_cell(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld,
                                int64_t r, int64_t c, int64_t* nr, int64_t* nc, int* nv) {
  int v = fdr[r * ld + c];
  if (v >= 8) return 0;
  int64_t rr = r + ORC_DY[v], cc = c + ORC_DX[v];
  if (rr < 0 || rr >= rows || cc < 0 || cc >= cols) return 0;
  *nr = rr;
  *nc = cc;
  *nv = fdr[rr * ld + cc];
  return 1;
}

/* ---------------------------------------------------------------------------
 * single_tile_flow_accumulation, accumulation part --
 * src/overflow/flow_accumulation.py:95-144 (Barnes 2016, Alg. 1).
 * Same three passes; the FIFO is an index array with a head pointer instead of
 * list.pop(0) (the reference's pop(0) makes it O(N^2) but does not change the
 * dequeue order).  NODATA cells start at -9999 (:119-121), have inflow 0, are
 * enqueued (:129-132) and incremented (:137) -> -9998, exactly as the reference.
 * Returns 0, or -1 if the queue could not be allocated.
 * ------------------------------------------------------------------------- */
int orc_flow_accumulation(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld,
                          int64_t* fac, int64_t ld_fac) {
  int64_t n = rows * cols;
  if (n == 0) return 0;
  uint8_t* inflow = (uint8_t*)calloc((size_t)n, 1);
  int64_t* queue = (int64_t*)malloc((size_t)n * sizeof(int64_t));
  if (!inflow || !queue) {
    free(inflow);
    free(queue);
    return -1;
  }
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) fac[r * ld_fac + c] = 0;
  /* pass 1 (:116-124) */
  for (int64_t r = 0; r < rows; ++r) {
    for (int64_t c = 0; c < cols; ++c) {
      int v = fdr[r * ld + c];
      int64_t nr, nc;
      int nv;
      int inside = orc_next_cell(fdr, rows, cols, ld, r, c, &nr, &nc, &nv);
      if (v == ORC_DIR_NODATA) {
        fac[r * ld_fac + c] = ORC_FAC_NODATA;
        continue;
      }
      if (!inside || nv == ORC_DIR_NODATA) continue;
      inflow[nr * cols + nc] += 1;
    }
  }
  /* pass 2 (:129-132) */
  int64_t head = 0, tail = 0;
  for (int64_t i = 0; i < n; ++i)
    if (inflow[i] == 0) queue[tail++] = i;
  /* pass 3 (:135-144) */
  while (head < tail) {
    int64_t i = queue[head++];
    int64_t r = i / cols, c = i % cols;
    fac[r * ld_fac + c] += 1;
    int64_t nr, nc;
    int nv;
    if (!orc_next_cell(fdr, rows, cols, ld, r, c, &nr, &nc, &nv) || nv == ORC_DIR_NODATA)
      continue;
    fac[nr * ld_fac + nc] += fac[r * ld_fac + c];
    if (--inflow[nr * cols + nc] == 0) queue[tail++] = nr * cols + nc;
  }
  free(inflow);
  free(queue);
  return 0;
}

/* ---------------------------------------------------------------------------
 * follow_path -- src/overflow/flow_accumulation.py:54-92 (Barnes 2016, Alg. 2)
 * out[0], out[1] = FLOW_EXTERNAL (-2,-2) / FLOW_TERMINATES (-1,-1) / exit (row,col).
 * The reference has no cycle guard (`while True`); max_steps bounds the walk and
 * a walk that exceeds it returns -1 (never happens on flow_direction output).
 * ------------------------------------------------------------------------- */
static int orc_follow_path(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld,
                           int64_t r0, int64_t c0, int64_t max_steps, int64_t* out) {
  int64_t r = r0, c = c0;
  for (int64_t step = 0; step <= max_steps; ++step) {
    int64_t nr, nc;
    int nv;
    if (!orc_next_cell(fdr, rows, cols, ld, r, c, &nr, &nc, &nv)) {
      if (r == r0 && c == c0) {
        out[0] = -2; /* FLOW_EXTERNAL, constants.py:59 */
        out[1] = -2;
      } else {
        out[0] = r;
        out[1] = c;
      }
      return 0;
    }
    if (nv == ORC_DIR_NODATA || nv == ORC_DIR_UNDEFINED) {
      out[0] = -1; /* FLOW_TERMINATES, constants.py:58 */
      out[1] = -1;
      return 0;
    }
    r = nr;
    c = nc;
  }
  return -1;
}

/* ---------------------------------------------------------------------------
 * perimeter_indices + links part of single_tile_flow_accumulation --
 * src/overflow/flow_accumulation.py:40-51,146-157.
 * perim_links is [n_perim][2] in perimeter_indices order: for every row the left
 * then the right column, then for cols 1..C-2 the top then the bottom row
 * (duplicates included when rows==1 or cols==1, as in the reference list).
 * perim_rc (nullable) receives the (row,col) of each entry.
 * Returns the number of entries written, or -1 on a runaway walk.
 * ------------------------------------------------------------------------- */
int64_t orc_perimeter_count(int64_t rows, int64_t cols) {
  int64_t inner = cols - 2 > 0 ? cols - 2 : 0;
  return 2 * rows + 2 * inner;
}

int64_t orc_links_perimeter(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld,
                            int64_t* perim_links, int64_t* perim_rc) {
  int64_t k = 0;
  int64_t max_steps = rows * cols;
  for (int64_t i = 0; i < rows; ++i) {
    int64_t cs[2] = {0, cols - 1};
    for (int j = 0; j < 2; ++j) {
      if (perim_rc) {
        perim_rc[2 * k] = i;
        perim_rc[2 * k + 1] = cs[j];
      }
      if (orc_follow_path(fdr, rows, cols, ld, i, cs[j], max_steps, perim_links + 2 * k)) return -1;
      ++k;
    }
  }
  for (int64_t j = 1; j < cols - 1; ++j) {
    int64_t rs[2] = {0, rows - 1};
    for (int i = 0; i < 2; ++i) {
      if (perim_rc) {
        perim_rc[2 * k] = rs[i];
        perim_rc[2 * k + 1] = j;
      }
      if (orc_follow_path(fdr, rows, cols, ld, rs[i], j, max_steps, perim_links + 2 * k)) return -1;
      ++k;
    }
  }
  return k;
}

/* ---------------------------------------------------------------------------
 * Size-independent exactness check: on an acyclic D8 graph
 * fac is the unique solution of
 *     fac[c] = 1 + sum of fac[u] over cells u whose downstream cell is c   (data)
 *     fac[c] = -9998                                                       (nodata)
 * with the edge rule of flow_accumulation.py:116-124 (u has code 0..7, c is
 * inside the raster and c is not NODATA).  Returns the number of violating cells.
 * ------------------------------------------------------------------------- */
int64_t orc_check_accumulation(const uint8_t* fdr, int64_t rows, int64_t cols, int64_t ld,
                               const int64_t* fac, int64_t ld_fac) {
  int64_t bad = 0;
#pragma omp parallel for schedule(static) reduction(+ : bad)
  for (int64_t r = 0; r < rows; ++r) {
    for (int64_t c = 0; c < cols; ++c) {
      int v = fdr[r * ld + c];
      int64_t want;
      if (v == ORC_DIR_NODATA) {
        want = ORC_FAC_NODATA + 1;
      } else {
        want = 1;
        for (int i = 0; i < 8; ++i) {
          int64_t ur = r + ORC_DY[i], uc = c + ORC_DX[i];
          if (ur < 0 || ur >= rows || uc < 0 || uc >= cols) continue;
          /* neighbour in direction i flows into (r,c) iff its code is the opposite */
          if (fdr[ur * ld + uc] == ((i + 4) & 7)) want += fac[ur * ld_fac + uc];
        }
      }
      if (fac[r * ld_fac + c] != want) ++bad;
    }
  }
  return bad;
}
