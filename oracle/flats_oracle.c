/*
 * flats_oracle.c -- CPU restatement of overflow's flat resolution (src/overflow/fix_flats.py,
 * Barnes, Lehman & Mulla 2014), SURVEY.md section 8(f) rank 2.
 *
 * TEST INFRASTRUCTURE ONLY (see d8_oracle.c): the parity oracle for overflow_b200/csrc/flats.cu.
 * Nothing under overflow_b200/ links or calls it.
 *
 * Parity status: PINNED.  oracle/gen_golden_flats.py runs the reference's own numba functions
 * (flat_edges, resolve_flats, d8_masked_flow_dirs, and the known-answer fixtures of
 * tests/test_fix_flats.py) and stores their outputs in tests/golden/fix_flats.npz;
 * tests/test_oracle_flats.py checks every function below against them.
 *
 * The restatement keeps the reference's order of operations (row-major scans, FIFO queues with a
 * level marker, labels numbered in the order low edges are met); the only change is an O(1) FIFO
 * (array + head index) where the reference uses list.pop(0) (fix_flats.py:96,147,203), which keeps
 * the dequeue order.  Paths below are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_DIR_UNDEFINED 8 /* src/overflow/constants.py:23 */
#define ORC_DIR_NODATA 9    /* src/overflow/constants.py:24 */

/* scan order E, NE, N, NW, W, SW, S, SE -- src/overflow/constants.py:29-40 */
static const int FL_DY[8] = {0, -1, -1, -1, 0, 1, 1, 1};
static const int FL_DX[8] = {1, 1, 0, -1, -1, -1, 0, 1};

typedef struct {
  int64_t* v;
  int64_t head, tail, cap;
} fifo_t;

static int fifo_push(fifo_t* q, int64_t x) {
  if (q->tail == q->cap) {
    int64_t ncap = q->cap ? q->cap * 2 : 1024;
    int64_t* nv = (int64_t*)realloc(q->v, (size_t)ncap * sizeof(int64_t));
    if (!nv) return -1;
    q->v = nv;
    q->cap = ncap;
  }
  q->v[q->tail++] = x;
  return 0;
}

/*
 * flat_edges -- src/overflow/fix_flats.py:13-62.
 * edges[r*cols+c]: bit 0 = low edge, bit 1 = high edge.  The reference appends to two lists while
 * scanning row-major; a list is recovered from the flags by a row-major scan.  Returns the number of
 * low edges; *n_high receives the number of high edges.
 */
int64_t orc_flat_edges_f32(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, uint8_t* edges,
                           int64_t* n_high) {
  int64_t n_lo = 0, n_hi = 0;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      uint8_t flag = 0;
      const uint8_t cur = fdr[r * cols + c];
      const float z = dem[r * cols + c];
      for (int k = 0; k < 8; ++k) { /* neighbor_generator, util/raster.py:239-272: in-bounds only */
        const int64_t nr = r + FL_DY[k], nc = c + FL_DX[k];
        if (nr < 0 || nr >= rows || nc < 0 || nc >= cols) continue;
        const uint8_t fn = fdr[nr * cols + nc];
        if (fn == ORC_DIR_NODATA) continue; /* :41-43 */
        const float zn = dem[nr * cols + nc];
        if (cur != ORC_DIR_UNDEFINED && fn == ORC_DIR_UNDEFINED && z == zn) { /* :45-53 */
          flag = 1;
          ++n_lo;
          break;
        }
        if (cur == ORC_DIR_UNDEFINED && z < zn) { /* :54-60 */
          flag = 2;
          ++n_hi;
          break;
        }
      }
      edges[r * cols + c] = flag;
    }
  if (n_high) *n_high = n_hi;
  return n_lo;
}

/* label_flats -- fix_flats.py:65-108: flood fill over cells of the seed's elevation, 8-connected. */
static int label_flats(const float* dem, int32_t* labels, int32_t new_label, int64_t r0, int64_t c0, int64_t rows,
                       int64_t cols, fifo_t* q) {
  q->head = q->tail = 0;
  const float elev = dem[r0 * cols + c0];
  if (fifo_push(q, r0 * (cols + 2) + c0 + 1 + (cols + 2)) != 0) return -1; /* padded coordinates: may leave the raster */
  while (q->head < q->tail) {
    const int64_t p = q->v[q->head++];
    const int64_t r = p / (cols + 2) - 1, c = p % (cols + 2) - 1;
    if (r < 0 || r >= rows || c < 0 || c >= cols) continue; /* :97-99 */
    if (dem[r * cols + c] != elev) continue;                /* :100-101 */
    if (labels[r * cols + c] != 0) continue;                /* :102-103 */
    labels[r * cols + c] = new_label;
    for (int k = 0; k < 8; ++k) /* :106-109: all eight, bounds checked at pop */
      if (fifo_push(q, (r + FL_DY[k] + 1) * (cols + 2) + (c + FL_DX[k] + 1)) != 0) return -1;
  }
  return 0;
}

/*
 * away_from_higher (fix_flats.py:111-161) when towards == 0, towards_lower (:164-224) when 1.
 * `seed` is the edge list in list order; the level marker is the value -1.
 */
static int flat_gradient(const int32_t* labels, int32_t* flat_mask, const uint8_t* fdr, const int64_t* seed,
                         int64_t n_seed, int32_t* flat_height, int64_t rows, int64_t cols, int towards, fifo_t* q) {
  q->head = q->tail = 0;
  if (towards)
    for (int64_t i = 0; i < rows * cols; ++i) flat_mask[i] = -flat_mask[i]; /* :200 */
  for (int64_t i = 0; i < n_seed; ++i)
    if (fifo_push(q, seed[i]) != 0) return -1;
  int32_t loops = 1;
  if (fifo_push(q, -1) != 0) return -1;
  while (q->tail - q->head > 1) { /* :145 / :202 */
    const int64_t p = q->v[q->head++];
    if (p < 0) {
      ++loops;
      if (fifo_push(q, -1) != 0) return -1;
      continue;
    }
    if (flat_mask[p] > 0) continue;
    if (!towards) {
      flat_mask[p] = loops; /* :153-154 */
      flat_height[labels[p] - 1] = loops;
    } else if (flat_mask[p] < 0) {
      flat_mask[p] += flat_height[labels[p] - 1] + 2 * loops; /* :211-212 */
    } else {
      flat_mask[p] = 2 * loops; /* :213-214 */
    }
    const int64_t r = p / cols, c = p % cols;
    for (int k = 0; k < 8; ++k) {
      const int64_t nr = r + FL_DY[k], nc = c + FL_DX[k];
      if (nr < 0 || nr >= rows || nc < 0 || nc >= cols) continue;
      const int64_t n = nr * cols + nc;
      if (labels[n] == labels[p] && fdr[n] == ORC_DIR_UNDEFINED) /* :155-161 / :215-224 */
        if (fifo_push(q, n) != 0) return -1;
    }
  }
  return 0;
}

/*
 * resolve_flats -- fix_flats.py:227-288.  flat_mask and labels are int32 [rows*cols] outputs.
 * Returns the number of labels handed out (>= 0) or -1 when out of memory.
 */
int64_t orc_resolve_flats_f32(const float* dem, const uint8_t* fdr, int64_t rows, int64_t cols, int32_t* flat_mask,
                              int32_t* labels) {
  const int64_t n = rows * cols;
  memset(flat_mask, 0, (size_t)n * sizeof(int32_t));
  memset(labels, 0, (size_t)n * sizeof(int32_t));
  if (n == 0) return 0;
  uint8_t* edges = (uint8_t*)malloc((size_t)n);
  if (!edges) return -1;
  int64_t n_hi = 0;
  const int64_t n_lo = orc_flat_edges_f32(dem, fdr, rows, cols, edges, &n_hi);
  if (n_lo == 0) { /* :258-264 */
    free(edges);
    return 0;
  }
  fifo_t q = {0, 0, 0, 0};
  int64_t* lo = (int64_t*)malloc((size_t)n_lo * sizeof(int64_t));
  int64_t* hi = (int64_t*)malloc((size_t)(n_hi ? n_hi : 1) * sizeof(int64_t));
  int32_t* flat_height = NULL;
  int64_t rc = -1;
  if (!lo || !hi) goto done;
  {
    int64_t a = 0, b = 0;
    for (int64_t i = 0; i < n; ++i) {
      if (edges[i] & 1) lo[a++] = i;
      if (edges[i] & 2) hi[b++] = i;
    }
  }
  int32_t label = 1;
  for (int64_t i = 0; i < n_lo; ++i) /* :266-271 */
    if (labels[lo[i]] == 0) {
      if (label_flats(dem, labels, label, lo[i] / cols, lo[i] % cols, rows, cols, &q) != 0) goto done;
      ++label;
    }
  {
    int64_t b = 0; /* :273-274 */
    for (int64_t i = 0; i < n_hi; ++i)
      if (labels[hi[i]] != 0) hi[b++] = hi[i];
    n_hi = b;
  }
  flat_height = (int32_t*)calloc((size_t)label, sizeof(int32_t)); /* :279-280 */
  if (!flat_height) goto done;
  if (flat_gradient(labels, flat_mask, fdr, hi, n_hi, flat_height, rows, cols, 0, &q) != 0) goto done;
  if (flat_gradient(labels, flat_mask, fdr, lo, n_lo, flat_height, rows, cols, 1, &q) != 0) goto done;
  rc = label - 1;
done:
  free(flat_height);
  free(hi);
  free(lo);
  free(q.v);
  free(edges);
  return rc;
}

/*
 * d8_masked_flow_dirs -- fix_flats.py:291-339.  Rewrites the UNDEFINED cells of fdr in place; the
 * scan reads only flat_mask and labels, so its result does not depend on the visiting order.
 */
void orc_d8_masked_flow_dirs(const int32_t* flat_mask, uint8_t* fdr, const int32_t* labels, int64_t rows,
                             int64_t cols) {
  const double sqrt2 = sqrt(2.0);
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < cols; ++c) {
      const int64_t i = r * cols + c;
      if (fdr[i] != ORC_DIR_UNDEFINED) continue; /* :317-320 (NODATA is != UNDEFINED too) */
      uint8_t nmin = ORC_DIR_UNDEFINED;
      double min_slope = INFINITY;
      for (int k = 0; k < 8; ++k) {
        const int64_t nr = r + FL_DY[k], nc = c + FL_DX[k];
        if (nr < 0 || nr >= rows || nc < 0 || nc >= cols) continue;
        const int64_t n = nr * cols + nc;
        if (labels[n] != labels[i]) continue;                           /* :332-333 */
        const double dz = (double)flat_mask[n] - (double)flat_mask[i];  /* :335 */
        const double slope = dz / ((FL_DY[k] != 0 && FL_DX[k] != 0) ? sqrt2 : 1.0);
        if (slope < min_slope) { /* :338-340 */
          min_slope = slope;
          nmin = (uint8_t)k;
        }
      }
      fdr[i] = nmin;
    }
}
