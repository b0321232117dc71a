"""Stand-in strip engine for CPU tests of overflow_b200.strips (TEST INFRASTRUCTURE).

Same interface as CudaStripEngine, computed with the CPU oracle and plain numpy/Python: an independent
restatement of the strip decomposition (Barnes 2016 at the strip level) used to check the host-side
exchange logic under gloo, where no GPU exists.
"""
import numpy as np
import torch

import oracle

DY = [0, -1, -1, -1, 0, 1, 1, 1]
DX = [1, 1, 0, -1, -1, -1, 0, 1]


class NumpyStripEngine:
    def empty(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def strip_workspace(self, rows, cols):
        return torch.zeros(1, dtype=torch.uint8)

    def boundary_workspace(self, n, cols):
        return torch.zeros(1, dtype=torch.uint8)

    def direction(self, dem_halo, nodata, fdr_out):
        dem = dem_halo.numpy()
        pad = np.full((dem.shape[0], dem.shape[1] + 2), np.float32(nodata), dtype=np.float32)
        pad[:, 1:-1] = dem
        fdr_out.copy_(torch.from_numpy(oracle.flow_direction_for_tile(pad, nodata)[1:-1, 1:-1].copy()))

    def flags(self, ws, rows, cols, bws, n_strips, out):
        out.zero_()

    def check(self, fdr_halo, fac, fac_above, fac_below):
        """Recurrence violations of the strip, its boundary rows checked against the neighbours' counts."""
        fh, f = fdr_halo.numpy(), fac.numpy()
        h, c = f.shape
        ext = np.zeros((h + 2, c), dtype=np.int64)
        ext[1:-1] = f
        codes = fh.copy()
        for row, vals in ((0, fac_above), (h + 1, fac_below)):
            if vals is None:
                codes[row] = 9  # nothing flows in from beyond the raster
            else:
                ext[row] = vals.numpy()
        bad = 0
        for y in range(1, h + 1):
            for x in range(c):
                if codes[y, x] == 9:
                    want = -9998
                else:
                    want = 1
                    for d in range(8):
                        uy, ux = y + DY[d], x + DX[d]
                        if 0 <= ux < c and codes[uy, ux] == (d + 4) % 8:
                            want += ext[uy, ux]
                bad += int(ext[y, x] != want)
        return bad

    def accum_local(self, fdr_halo, has_above, has_below, fac, ws, rec):
        from overflow_b200.strips import record_views

        floc, slink, bcode = record_views(rec, fac.shape[1])
        fh = fdr_halo.numpy()
        fdr = np.ascontiguousarray(fh[1:-1])
        h, c = fdr.shape
        local = oracle.flow_accumulation(fdr)  # edges leaving the strip are dropped: strip-local counts
        fac.copy_(torch.from_numpy(local))
        rc, links = oracle.links_perimeter(fdr)
        lk = {(int(r), int(cc)): (int(a), int(b)) for (r, cc), (a, b) in zip(rc, links)}
        for t, r in enumerate((0, h - 1)):
            for x in range(c):
                bcode[t, x] = int(fdr[r, x])
                floc[t, x] = int(local[r, x])
                sl = -1
                if fdr[r, x] < 8:
                    a, b = lk[(r, x)]
                    if (a, b) == (-2, -2):
                        a, b = r, x
                    if a >= 0:
                        code = int(fdr[a, b])
                        ny, nx = a + DY[code], b + DX[code]
                        if 0 <= nx < c and ((ny == -1 and has_above) or (ny == h and has_below)):
                            if fh[ny + 1, nx] != 9:
                                sl = ((0 if a == 0 else 1) << 30) | b
                slink[t, x] = sl

    def boundary_solve(self, rec_all, cols, J_all, ws):
        from overflow_b200.strips import record_views

        floc_all, slink_all, bcode_all = record_views(rec_all, cols)
        sl, fl, bc = slink_all.numpy(), floc_all.numpy(), bcode_all.numpy()
        G, _, C = sl.shape

        def target(s, t, c):
            code = int(bc[s, t, c])
            if code >= 8:
                return None
            dy = DY[code]
            if (t == 0 and dy != -1) or (t == 1 and dy != 1):
                return None
            s2, c2 = s + dy, c + DX[code]
            if not (0 <= s2 < G and 0 <= c2 < C):
                return None
            t2 = 1 if dy < 0 else 0
            return None if bc[s2, t2, c2] == 9 else (s2, t2, c2)

        base = np.zeros((G, 2, C), dtype=np.int64)
        parent = {}
        for s in range(G):
            for t in range(2):
                for c in range(C):
                    d = target(s, t, c)
                    if d is not None:
                        base[d] += fl[s, t, c]
                    v = int(sl[s, t, c])
                    if v >= 0:
                        d2 = target(s, v >> 30, v & ((1 << 30) - 1))
                        if d2 is not None:
                            parent[(s, t, c)] = d2
        J = base.copy()
        # subtree sums: every node adds its base to each of its ancestors
        for s in range(G):
            for t in range(2):
                for c in range(C):
                    if base[s, t, c] == 0:
                        continue
                    n, steps = (s, t, c), 0
                    while n in parent:
                        n = parent[n]
                        J[n] += base[s, t, c]
                        steps += 1
                        assert steps <= 2 * G * C
        J_all.copy_(torch.from_numpy(J))

    def accum_final(self, fdr_halo, has_above, has_below, J_mine, ws, fac):
        fdr = fdr_halo.numpy()[1:-1]
        h, c = fdr.shape
        out = oracle.flow_accumulation(np.ascontiguousarray(fdr))
        J = J_mine.numpy()
        for t, r in enumerate((0, h - 1)):
            for x in range(c):
                j = int(J[t, x])
                if j == 0:
                    continue
                y, xx = r, x
                while True:
                    out[y, xx] += j
                    code = int(fdr[y, xx])
                    if code >= 8:
                        break
                    ny, nx = y + DY[code], xx + DX[code]
                    if not (0 <= ny < h and 0 <= nx < c) or fdr[ny, nx] == 9:
                        break
                    y, xx = ny, nx
        fac.copy_(torch.from_numpy(out))
