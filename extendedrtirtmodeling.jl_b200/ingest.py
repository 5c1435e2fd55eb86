"""Ingest side of the boundary: a CSV table -> InputData, the step the reference leaves to CSV.jl +
DataFrames in user code (README.md:85-96: Y = columns of 0/1 responses, T = response times in seconds, X = covariates).
The engine itself takes the resulting column-major float64 arrays (erirt_set_data) and re-lays them out on the device."""
import numpy as np

from .structs import InputData


def _select(header, cols):
    if isinstance(cols, slice):
        return list(range(len(header)))[cols]
    out = []
    for c in cols:
        out.append(header.index(c) if isinstance(c, str) else int(c))
    return out


def readCsvData(path, y_cols, t_cols=None, x_cols=None, t_is_log=False, delimiter=",", drop_missing=True):
    """Read a delimited text file with one header row.  y_cols / t_cols / x_cols: lists of column names or 0-based indices
    (or slices).  Response times are taken as seconds (InputData computes logT = log.(T), src/Base.pl.jl:73-77) unless
    t_is_log.  Rows with a missing / non-numeric entry in a selected column are dropped when drop_missing (the reference has
    no missing-data handling: a `missing` in Y or T makes its constructors throw), otherwise a ValueError names the first one."""
    with open(path, "r", encoding="utf-8-sig") as fh:
        header = [h.strip().strip('"') for h in fh.readline().rstrip("\r\n").split(delimiter)]
        yi = _select(header, y_cols)
        ti = _select(header, t_cols) if t_cols is not None else []
        xi = _select(header, x_cols) if x_cols is not None else []
        want = yi + ti + xi
        rows, bad = [], []
        for ln, line in enumerate(fh, start=2):
            if not line.strip():
                continue
            f = line.rstrip("\r\n").split(delimiter)
            try:
                vals = [float(f[i]) for i in want]
            except (ValueError, IndexError):
                bad.append(ln)
                continue
            if not all(np.isfinite(vals)):
                bad.append(ln)
                continue
            rows.append(vals)
    if bad and not drop_missing:
        raise ValueError(f"{path}: missing or non-numeric value in a selected column at line {bad[0]} ({len(bad)} such rows)")
    if not rows:
        raise ValueError(f"{path}: no complete rows")
    A = np.asarray(rows, dtype=np.float64)
    Y = A[:, : len(yi)]
    if not np.all((Y == 0) | (Y == 1)):
        raise ValueError("response columns must hold 0/1")
    T = A[:, len(yi): len(yi) + len(ti)] if ti else None
    if T is not None and t_is_log:
        T = np.exp(T)
    if T is not None and not np.all(T > 0):
        raise ValueError("response times must be positive")
    X = A[:, len(yi) + len(ti):] if xi else None
    D = InputData(Y=Y, T=T, X=X)
    D.dropped_rows = bad
    return D
