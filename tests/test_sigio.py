"""Sourmash-free signature reading against the reference's own fixture
(reference tests/test_unittests.py:144-156 and tests/unittests_data/test_collect_signature_info_data.json)."""
import json
import os

import numpy as np
import pytest

from yacht_b200 import sigio

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_roundtrip_and_md5(tmp_path):
    z = np.load(os.path.join(GOLD, "fixture20.npz"))
    off = z["offsets"]
    for g in (0, 7, 19):
        mins = z["hashes"][int(off[g]):int(off[g + 1])]
        # the md5 sourmash stored in the fixture is reproduced from (ksize, mins)
        assert sigio.compute_md5sum(31, mins) == str(z["md5"][g])
        p = str(tmp_path / f"{g}.sig.gz")
        sigio.write_signature(p, str(z["names"][g]), mins, 31, abundances=np.ones(len(mins), dtype=np.int64))
        s = sigio.load_signature_with_ksize(p, 31)
        assert np.array_equal(s.mins, mins) and s.name == str(z["names"][g]) and s.scaled == 1000
        assert s.md5sum == str(z["md5"][g]) and s.mean_abundance == 1.0


def test_sig_info_matches_reference_fixture():
    path = os.path.join(GOLD, "collect_signature_info_expected.json")
    with open(path) as f:
        expected = json.load(f)
    z = np.load(os.path.join(GOLD, "fixture20.npz"))
    off = z["offsets"]
    abund = z["abundances"]
    assert len(expected) == 20
    for g in range(20):
        name = str(z["names"][g])
        md5, mean_ab, n_hashes, scaled = expected[name]
        assert md5 == str(z["md5"][g])
        assert n_hashes == int(off[g + 1] - off[g])
        assert scaled == int(round((2 ** 64 - 1) / int(z["max_hash"][g])))
        ab = abund[int(off[g]):int(off[g + 1])]
        assert abs(float(np.mean(ab)) - mean_ab) < 1e-12


def test_errors(tmp_path):
    p = str(tmp_path / "empty.sig")
    sigio.write_signature(p, "empty", [], 31)
    # reference tests/test_utils_for_bug_YAC-13.py: an empty sketch raises this ValueError
    with pytest.raises(ValueError, match="Empty sketch in signature"):
        sigio.load_signature_with_ksize(p, 31)
    p2 = str(tmp_path / "a.sig")
    sigio.write_signature(p2, "a", [1, 2, 3], 21)
    with pytest.raises(ValueError, match="Expected exactly one signature with ksize 31"):
        sigio.load_signature_with_ksize(p2, 31)


def test_zip_database(tmp_path):
    zp = str(tmp_path / "db.zip")
    sk = [dict(name=f"g{k}", mins=[k + 1, k + 100, k + 1000], abundances=[1, 2, 3]) for k in range(4)]
    sigio.write_sig_zip(zp, sk, 31)
    back = sigio.read_sig_zip(zp)
    assert [s.name for s in back] == ["g0", "g1", "g2", "g3"]
    assert back[2].mean_abundance == 2.0 and list(back[2].mins) == [3, 102, 1002]


def test_write_signatures_round_trip(tmp_path):
    """several sketches in one JSON signature file (what `sourmash sketch dna -o x.sig` writes for several records)"""
    import numpy as np
    from yacht_b200 import sigio
    sketches = [dict(name="rec one", filename="a.fa", mins=np.array([3, 17, 99], np.uint64), abundances=np.array([1, 4, 2], np.uint32)),
                dict(name="", filename="a.fa", mins=np.array([5], np.uint64), abundances=None),
                dict(name="empty", filename="a.fa", mins=np.zeros(0, np.uint64), abundances=np.zeros(0, np.uint32))]
    for fn in ("many.sig", "many.sig.gz"):
        path = str(tmp_path / fn)
        sigio.write_signatures(path, sketches, ksize=21, max_hash=sigio.MAX_HASH_SCALED_1000)
        back = sigio.parse_signature_json(sigio._open_text(path), path)
        assert [s.name for s in back] == ["rec one", "", "empty"]
        assert [list(map(int, s.mins)) for s in back] == [[3, 17, 99], [5], []]
        assert list(map(int, back[0].abundances)) == [1, 4, 2] and back[1].abundances is None
        assert all(s.ksize == 21 and s.scaled == 1000 for s in back)
        assert back[0].md5sum == sigio.compute_md5sum(21, [3, 17, 99])
