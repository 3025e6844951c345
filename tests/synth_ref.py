"""numpy restatement of the device generators in smvp-toolkit_b200/csrc/synth.cu (test infrastructure)."""
import numpy as np

from oracle import oracle

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & M64
        return x ^ (x >> np.uint64(31))


def hash_uniform(h):
    return 2.0 * ((h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)) - 1.0


def hash_value(seed, row, col):
    key = (np.asarray(row, np.uint64) << np.uint64(32)) | np.asarray(col, np.uint64)
    return hash_uniform(splitmix64(np.uint64(seed) ^ splitmix64(key)))


def vector(n, seed):
    if seed == 0:
        return np.ones(n)
    return hash_uniform(splitmix64(np.uint64(seed) ^ splitmix64(np.arange(n, dtype=np.uint64))))


def stencil27(nx, ny, nz, row_begin=0, row_end=None, value_mode=0, seed=0):
    """COO (oracle.COO_DT) of rows [row_begin,row_end), local row indices, (row,col)-sorted."""
    total = nx * ny * nz
    row_end = total if row_end is None else row_end
    rows, cols, vals = [], [], []
    for r in range(row_begin, row_end):
        ix, iy, iz = r % nx, (r // nx) % ny, r // (nx * ny)
        for dz in (-1, 0, 1):
            z = iz + dz
            if z < 0 or z >= nz:
                continue
            for dy in (-1, 0, 1):
                y = iy + dy
                if y < 0 or y >= ny:
                    continue
                for dx in (-1, 0, 1):
                    x = ix + dx
                    if x < 0 or x >= nx:
                        continue
                    c = x + nx * (y + ny * z)
                    rows.append(r - row_begin)
                    cols.append(c)
                    vals.append((26.0 if c == r else -1.0) if value_mode == 0 else 0.0)
    coo = oracle.make_coo(np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals, np.float64))
    if value_mode == 1:
        coo["val"] = hash_value(seed, coo["row"].astype(np.int64) + row_begin, coo["col"])
    elif value_mode == 2:
        coo["val"] = 1.0
    return coo


def stencil27_row_counts(nx, ny, nz):
    def span(n):
        s = np.full(n, 3)
        s[0] -= 1
        s[-1] -= 1
        return s if n > 1 else np.ones(1, int)
    return (span(nz)[:, None, None] * span(ny)[None, :, None] * span(nx)[None, None, :]).reshape(-1)


def rmat(scale, nedges, a=0.57, b=0.19, c=0.19, value_mode=1, seed=42):
    e = np.arange(nedges, dtype=np.uint64)
    base = splitmix64(np.uint64(seed) ^ e)
    r = np.zeros(nedges, np.uint64)
    cc = np.zeros(nedges, np.uint64)
    ab, abc = a + b, a + b + c
    for lvl in range(scale):
        with np.errstate(over="ignore"):
            u = (splitmix64((base + np.uint64(lvl)) & M64) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
        rb = (u >= ab).astype(np.uint64)
        cb = (((u >= a) & (u < ab)) | (u >= abc)).astype(np.uint64)
        r = (r << np.uint64(1)) | rb
        cc = (cc << np.uint64(1)) | cb
    key = np.unique((r << np.uint64(scale)) | cc)
    row = (key >> np.uint64(scale)).astype(np.int32)
    col = (key & np.uint64((1 << scale) - 1)).astype(np.int32)
    val = hash_value(seed, row, col) if value_mode == 1 else np.ones(len(key))
    return oracle.make_coo(row, col, val)
