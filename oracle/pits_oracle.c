/*
 * pits_oracle.c -- CPU restatement of overflow's single-cell pit breaching
 * (src/overflow/breach_single_cell_pits.py:9-63), SURVEY.md section 8(f) rank 4.
 *
 * TEST INFRASTRUCTURE ONLY (see d8_oracle.c): the parity oracle for overflow_b200/csrc/pits.cu.
 *
 * Parity status: PINNED.  oracle/gen_golden_pits.py runs the reference's own numba function
 * (breach_single_cell_pits_in_chunk, including the fixture of tests/test_breach_single_cell_pits.py:38-83)
 * and stores its outputs in tests/golden/breach_pits.npz; tests/test_oracle_pits.py checks this file
 * against them bit for bit.
 *
 * Types as numba resolves them for a float32 chunk: z + zn is a float32 sum, the division by the
 * integer 2 happens in float64 and the store rounds back to float32 (:60); the nodata value is a
 * Python float, so `zn == nodata_value` compares in float64 (:43, :59).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static const int P_DX[8] = {1, 1, 1, 0, -1, -1, -1, 0};   /* :27 */
static const int P_DY[8] = {-1, 0, 1, 1, 1, 0, -1, -1};   /* :28 */
static const int P_DX2[16] = {2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1};  /* :29 */
static const int P_DY2[16] = {-2, -1, 0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2};  /* :30 */
static const int P_BREACH[16] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 0};      /* :31 */

/*
 * chunk: float32 rows x cols, modified in place; unsolved: int8 rows x cols out (1 = a single-cell pit that
 * could not be breached).  Returns the number of pits found by the first pass.
 */
int64_t orc_breach_single_cell_pits_f32(float* chunk, int64_t rows, int64_t cols, double nodata, int8_t* unsolved) {
  memset(unsolved, 0, (size_t)(rows * cols));
  int64_t n_pits = 0;
  /* :38-50 -- pits of the chunk as it came in (the reference runs this pass in parallel: nothing is written) */
  for (int64_t r = 2; r < rows - 2; ++r)
    for (int64_t c = 2; c < cols - 2; ++c) {
      const float z = chunk[r * cols + c];
      if ((double)z != nodata) {
        int flag = 1;
        for (int k = 0; k < 8; ++k) {
          const float zn = chunk[(r + P_DY[k]) * cols + c + P_DX[k]];
          if (zn <= z || (double)zn == nodata) {
            flag = 0;
            break;
          }
        }
        if (flag) {
          unsolved[r * cols + c] = 1;
          ++n_pits;
        }
      }
    }
  /* :52-63 -- the pits in row-major order (np.argwhere), each reading the chunk as the earlier ones left it */
  for (int64_t r = 2; r < rows - 2; ++r)
    for (int64_t c = 2; c < cols - 2; ++c) {
      if (!unsolved[r * cols + c]) continue;
      int solved = 0;
      const float z = chunk[r * cols + c];
      for (int k = 0; k < 16; ++k) {
        const float zn = chunk[(r + P_DY2[k]) * cols + c + P_DX2[k]];
        if (zn <= z || (double)zn == nodata) {
          solved = 1;
          const float sum = z + zn;
          chunk[(r + P_DY[P_BREACH[k]]) * cols + c + P_DX[P_BREACH[k]]] = (float)((double)sum / 2.0);
        }
      }
      if (solved) unsolved[r * cols + c] = 0;
    }
  return n_pits;
}
