"""CPU test of the CLI readers (demethify.py:103-143): the concurrent two-column reader must return exactly what the reference's
sequential pandas reads return, for bedmethyl and csv inputs, with and without --fillna, and for frequency-only csv files."""
import argparse

import numpy as np
import pandas as pd
import pytest

from demethify_b200.demethify import read_inputs


def reference_style(paths, ref_path, bedmethyl, fillna):
    """The reference's reader, restated (demethify.py:103-143)."""
    sep = "\t" if bedmethyl else ","
    ref = pd.read_csv(ref_path, sep=sep)
    if bedmethyl:
        ref = ref.iloc[:, 3:]
    if fillna:
        ref = ref.fillna(0)
    fr, cv = [], []
    for p in paths:
        t = pd.read_csv(p, sep=sep)
        if not bedmethyl and t.shape[1] == 1:
            t["valid_coverage"] = 1
        if fillna:
            t = t.fillna(0)
        fr.append(t["percent_modified"].values / 100 if bedmethyl else t["percent_modified"].values)
        cv.append(t["valid_coverage"].values)
    return np.column_stack(fr), np.column_stack(cv), ref.values, list(ref.columns)


@pytest.mark.parametrize("bedmethyl,fillna", [(True, False), (True, True), (False, False), (False, True)])
def test_readers_match_reference_style(tmp_path, bedmethyl, fillna):
    rs = np.random.RandomState(3)
    M, n = 257, 5
    sep = "\t" if bedmethyl else ","
    pos = pd.DataFrame({"chrom": ["chr1"] * M, "start": np.arange(M), "end": np.arange(M) + 1})
    refm = pd.DataFrame(rs.uniform(size=(M, 4)), columns=["A", "B", "C", "D"])
    if fillna:
        refm.iloc[5, 2] = np.nan
    ref_path = tmp_path / "ref"
    (pd.concat([pos, refm], axis=1) if bedmethyl else refm).to_csv(ref_path, sep=sep, index=False)
    paths = []
    for j in range(n):
        cov = rs.poisson(30, size=M) + 1
        cnt = rs.binomial(cov, rs.uniform(size=M))
        pm = cnt / cov * (100 if bedmethyl else 1)
        t = pd.DataFrame({"valid_coverage": cov, "count_modified": cnt, "percent_modified": pm})
        if fillna:
            t["percent_modified"] = t["percent_modified"].astype(float)
            t.loc[7 + j, "percent_modified"] = np.nan
        if bedmethyl:
            t = pd.concat([pos, t], axis=1)
        elif j == 2 and not fillna:
            t = t[["percent_modified"]]                    # csv with frequencies only -> coverage 1
        p = tmp_path / f"s{j}"
        t.to_csv(p, sep=sep, index=False)
        paths.append(str(p))
    args = argparse.Namespace(bedmethyl=bedmethyl, fillna=fillna, ref=str(ref_path), methfreq=paths)
    X, C, R, hdr = read_inputs(args)
    X0, C0, R0, hdr0 = reference_style(paths, str(ref_path), bedmethyl, fillna)
    assert hdr == hdr0 and X.shape == X0.shape and C.shape == C0.shape
    assert np.array_equal(X, X0, equal_nan=True) and np.array_equal(np.asarray(C, dtype=np.float64), np.asarray(C0, dtype=np.float64), equal_nan=True)
    assert np.array_equal(R, R0, equal_nan=True)
    assert X.flags.c_contiguous and C.flags.c_contiguous


def test_reader_rejects_ragged_files(tmp_path):
    a = tmp_path / "a.csv"; b = tmp_path / "b.csv"
    pd.DataFrame({"percent_modified": [0.1, 0.2, 0.3], "valid_coverage": [3, 4, 5]}).to_csv(a, index=False)
    pd.DataFrame({"percent_modified": [0.1, 0.2], "valid_coverage": [3, 4]}).to_csv(b, index=False)
    args = argparse.Namespace(bedmethyl=False, fillna=False, ref=None, methfreq=[str(a), str(b)])
    with pytest.raises(ValueError):
        read_inputs(args)


def test_reader_stages_integer_coverage_as_uint16_and_widens_when_needed(tmp_path):
    """Integer coverage lands in a uint16 matrix (what the kernels store); a file whose coverage does not fit 16 bits turns the
    whole matrix into int64 - values identical to the reference's column_stack either way."""
    rs = np.random.RandomState(1)
    M = 50
    paths = []
    for j, big in enumerate([False, False, True]):
        cov = rs.poisson(30, size=M) + 1
        if big:
            cov[7] = 70000
        p = tmp_path / f"w{j}.csv"
        pd.DataFrame({"percent_modified": rs.uniform(size=M), "valid_coverage": cov}).to_csv(p, index=False)
        paths.append(str(p))
    X, C, _, _ = read_inputs(argparse.Namespace(bedmethyl=False, fillna=False, ref=None, methfreq=paths[:2]))
    assert C.dtype == np.uint16
    X3, C3, _, _ = read_inputs(argparse.Namespace(bedmethyl=False, fillna=False, ref=None, methfreq=paths))
    assert C3.dtype == np.int64 and C3[7, 2] == 70000 and np.array_equal(C3[:, :2], C.astype(np.int64))
