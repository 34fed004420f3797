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


def test_quoted_fields_bom_and_signed_ids(tmp_path):
    """Exports of spreadsheet tools quote every field and start with a byte-order mark; both used to be read as an empty set
    (every line 'does not start with a digit'), and so was a line with a negative id. Quotes are separators now, the BOM is
    skipped, and a signed id is an error with its line number."""
    q = tmp_path / "quoted.csv"
    q.write_bytes(b'\xef\xbb\xbf"userId","movieId","rating","timestamp"\r\n"7","31","2.5","1260759144"\r\n"3","31","4.0","1260759179"\r\n')
    rf = mf.read_ratings(q)
    assert rf.format == capi.FORMAT_TRIPLETS and rf.userIds.tolist() == [3, 7] and rf.itemIds.tolist() == [31]
    assert rf.users.tolist() == [1, 0] and rf.ratings.tolist() == [2.5, 4.0]
    b = tmp_path / "bom_then_data.txt"
    b.write_bytes(b"\xef\xbb\xbf5 6 1.5\n8 6 3\n")
    rf = mf.read_ratings(b)
    assert rf.userIds.tolist() == [5, 8] and rf.ratings.tolist() == [1.5, 3.0]
    nb = tmp_path / "bom_netflix.txt"
    nb.write_bytes(b"\xef\xbb\xbf12:\n5,3,2005-01-01\n")
    rf = mf.read_ratings(nb)
    assert rf.format == capi.FORMAT_NETFLIX_PRIZE and rf.itemIds.tolist() == [12] and rf.userIds.tolist() == [5]
    for text in ("1 2 3\n-4 2 5\n", "1 2 3\n+4 2 5\n"):
        s = tmp_path / "signed.txt"
        s.write_text(text)
        with pytest.raises(mf.MfsgdError) as ei:
            mf.read_ratings(s)
        assert ei.value.code == capi.E_INVALID_ARG and "signed.txt:2" in str(ei.value) and "signed id" in str(ei.value)
    # a negative RATING is a value, not an id: kept
    n = tmp_path / "negrating.txt"
    n.write_text("1 2 -0.5\n")
    assert mf.read_ratings(n).ratings.tolist() == [-0.5]


@pytest.mark.parametrize("threads", [2, 3, 7, 16, 64, 200])
def test_netflix_prize_slices_find_their_movie(tmp_path, monkeypatch, threads):
    """Slice-parallel parsing: a slice that starts in the middle of a movie's block (or on an empty movie, or on a header line)
    looks backwards for the header it continues. MFSGD_IO_THREADS forces many slices on a small file."""
    rng = np.random.default_rng(threads)
    lines, want = [], []
    for m in range(1, 80):
        lines.append("%d:" % m)
        for _ in range(int(rng.integers(0, 25))):          # some movies have no ratings at all
            u, r = int(rng.integers(1, 1000)), int(rng.integers(1, 6))
            lines.append("%d,%d,2005-01-01" % (u, r))
            want.append((u, m, r))
    path = tmp_path / "combined.txt"
    path.write_text("\n".join(lines) + "\n")
    monkeypatch.setenv("MFSGD_IO_THREADS", str(threads))
    rf = mf.read_ratings(path)
    got = list(zip(rf.userIds[rf.users].tolist(), rf.itemIds[rf.items].tolist(), [int(x) for x in rf.ratings]))
    assert rf.format == capi.FORMAT_NETFLIX_PRIZE and got == want


@pytest.mark.parametrize("threads", [3, 16])
def test_triplet_slices_and_first_error_in_file_order(tmp_path, monkeypatch, threads):
    users, items, ratings = synth(3000, seed=11)
    path = tmp_path / "many.tsv"
    rows = ["%d\t%d\t%g" % (a, b, c) for a, b, c in zip(users, items, ratings)]
    path.write_text("\n".join(rows) + "\n")
    monkeypatch.setenv("MFSGD_IO_THREADS", str(threads))
    check(mf.read_ratings(path), users, items, ratings, capi.FORMAT_TRIPLETS)
    rows[2500] = "7\tx\t3"            # two bad lines in different slices: the first one in the file is reported
    rows[700] = "7\t\t"
    path.write_text("\n".join(rows) + "\n")
    with pytest.raises(mf.MfsgdError) as ei:
        mf.read_ratings(path)
    assert "many.tsv:701" in str(ei.value)
