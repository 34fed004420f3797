"""Ratings-file ingest (mfsgd_read_ratings): every supported text format against a plain-Python parse of
the same bytes. Host-only code: runs without a GPU."""
import numpy as np
import pytest

import matrixfactorizationsgd.java_b200 as mf
from matrixfactorizationsgd.java_b200 import _capi as capi


def synth(n=5000, seed=3):
    rng = np.random.default_rng(seed)
    users = rng.choice(np.array([7, 12, 13, 500, 90210, 1_000_003, 2_147_483_000]), n)     # sparse file ids
    items = rng.integers(1, 400, n) * 17
    ratings = rng.integers(1, 11, n) / 2.0                                                     # 0.5 .. 5.0
    return users, items, ratings


def expect(users, items, ratings):
    uid, u = np.unique(users, return_inverse=True)
    iid, i = np.unique(items, return_inverse=True)
    return u.astype(np.int32), i.astype(np.int32), ratings.astype(np.float32), uid.astype(np.int64), iid.astype(np.int64)


def check(rf, users, items, ratings, fmt):
    u, i, r, uid, iid = expect(users, items, ratings)
    assert rf.format == fmt and rf.nUsers == len(uid) and rf.nItems == len(iid)
    assert np.array_equal(rf.userIds, uid) and np.array_equal(rf.itemIds, iid)
    assert np.array_equal(rf.users, u) and np.array_equal(rf.items, i) and np.array_equal(rf.ratings, r)


@pytest.mark.parametrize("name,sep,header,tail", [
    ("u.data", "\t", "", "\t881250949"),                       # MovieLens-100K
    ("ratings.csv", ",", "userId,movieId,rating,timestamp\n", ",1147880044"),   # MovieLens-20M/25M
    ("ratings.dat", "::", "", "::978300760"),                  # MovieLens-1M/10M
    ("plain.txt", " ", "# user item rating\n\n", ""),
])
def test_triplet_formats(tmp_path, name, sep, header, tail):
    users, items, ratings = synth()
    path = tmp_path / name
    with open(path, "w") as f:
        f.write(header)
        for a, b, c in zip(users, items, ratings):
            f.write("%d%s%d%s%s%s\n" % (a, sep, b, sep, ("%g" % c) if name != "ratings.csv" else "%.1f" % c, tail))
    check(mf.read_ratings(path), users, items, ratings, capi.FORMAT_TRIPLETS)
    check(mf.read_ratings(path, capi.FORMAT_TRIPLETS), users, items, ratings, capi.FORMAT_TRIPLETS)


def test_netflix_prize_format(tmp_path):
    users, items, ratings = synth(4000, seed=5)
    order = np.argsort(items, kind="stable")
    users, items, ratings = users[order], items[order], np.round(ratings[order]).clip(1, 5)
    path = tmp_path / "combined_data_1.txt"
    with open(path, "w") as f:
        last = None
        for a, b, c in zip(users, items, ratings):
            if b != last:
                f.write("%d:\n" % b)
                last = b
            f.write("%d,%d,2005-09-06\n" % (a, int(c)))
    check(mf.read_ratings(path), users, items, ratings, capi.FORMAT_NETFLIX_PRIZE)


def test_no_trailing_newline_crlf_and_exponent(tmp_path):
    path = tmp_path / "odd.txt"
    path.write_bytes(b"3 4 2.5\r\n1,1,5e-1\r\n  9;9;4")
    rf = mf.read_ratings(path)
    assert rf.users.tolist() == [1, 0, 2] and rf.items.tolist() == [1, 0, 2]
    assert rf.ratings.tolist() == [2.5, 0.5, 4.0] and rf.userIds.tolist() == [1, 3, 9]


def test_empty_and_errors(tmp_path):
    empty = tmp_path / "empty.csv"
    empty.write_text("userId,movieId,rating\n")
    rf = mf.read_ratings(empty)
    assert len(rf.ratings) == 0 and rf.nUsers == 0 and rf.nItems == 0
    bad = tmp_path / "bad.txt"
    bad.write_text("1 2 3\n4 x 5\n")
    with pytest.raises(mf.MfsgdError) as ei:
        mf.read_ratings(bad)
    assert ei.value.code == capi.E_INVALID_ARG and "bad.txt:2" in str(ei.value)
    orphan = tmp_path / "orphan.txt"
    orphan.write_text("5,3,2005-01-01\n")
    with pytest.raises(mf.MfsgdError):
        mf.read_ratings(orphan, capi.FORMAT_NETFLIX_PRIZE)
    with pytest.raises(mf.MfsgdError) as ei:
        mf.read_ratings(tmp_path / "missing.txt")
    assert "cannot open" in str(ei.value)
    with pytest.raises(mf.MfsgdError):
        mf.read_ratings(empty, 9)


def test_large_sparse_ids_take_the_sort_path(tmp_path):
    path = tmp_path / "big_ids.txt"
    path.write_text("9000000000 5 1\n7 6000000000 2\n9000000000 6000000000 3\n")
    rf = mf.read_ratings(path)
    assert rf.userIds.tolist() == [7, 9000000000] and rf.itemIds.tolist() == [5, 6000000000]
    assert rf.users.tolist() == [1, 0, 1] and rf.items.tolist() == [0, 1, 1]
