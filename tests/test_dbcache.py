"""Packed sketch cache (yacht_b200/dbcache.py): round trip, keyed by the manifest's md5 list,
never touches the reference's own files, invalid or foreign caches are ignored."""
import json
import os

import numpy as np

from yacht_b200 import dbcache, synth


def _db(tmp_path):
    db = synth.make_reference_db(40, 5, mean_size=120, sd_size=30)
    md5 = [f"{g:032x}" for g in range(db.n)]
    d = tmp_path / "ref_intermediate_files"
    (d / "signatures").mkdir(parents=True)
    (d / "signatures" / "keep.sig").write_text("[]")
    return db, md5, str(d)


def test_round_trip_and_key(tmp_path):
    db, md5, d = _db(tmp_path)
    assert dbcache.load(d, md5) is None
    assert dbcache.store(d, md5, db.hashes, db.offsets)
    got = dbcache.load(d, md5)
    assert got is not None
    h, o = got
    assert np.array_equal(np.asarray(h), db.hashes) and np.array_equal(o, db.offsets.astype(np.uint64))
    assert dbcache.load(d, md5[:-1]) is None                    # another genome list
    assert dbcache.load(d, list(reversed(md5))) is None         # another ROW ORDER (row = genome id on the device)
    assert sorted(os.listdir(d)) == sorted([dbcache.DIRNAME, "signatures"])   # only a new sub-directory appeared
    assert os.listdir(os.path.join(d, "signatures")) == ["keep.sig"]


def test_corrupt_or_disabled_cache_is_ignored(tmp_path, monkeypatch):
    db, md5, d = _db(tmp_path)
    dbcache.store(d, md5, db.hashes, db.offsets)
    c = dbcache.cache_dir(d)
    np.save(os.path.join(c, "offsets.npy"), db.offsets.astype(np.uint64)[:-1])
    assert dbcache.load(d, md5) is None                         # offsets of the wrong length
    dbcache.store(d, md5, db.hashes, db.offsets)
    with open(os.path.join(c, "meta.json")) as f:
        meta = json.load(f)
    meta["version"] = 99
    with open(os.path.join(c, "meta.json"), "w") as f:
        json.dump(meta, f)
    assert dbcache.load(d, md5) is None
    dbcache.store(d, md5, db.hashes, db.offsets)
    os.remove(os.path.join(c, "hashes.npy"))
    assert dbcache.load(d, md5) is None
    dbcache.store(d, md5, db.hashes, db.offsets)
    monkeypatch.setenv("YACHT_DB_CACHE", "0")
    assert dbcache.load(d, md5) is None and not dbcache.store(d, md5, db.hashes, db.offsets)


def test_empty_database(tmp_path):
    d = str(tmp_path / "x")
    os.makedirs(d)
    assert dbcache.store(d, [], np.zeros(0, dtype=np.uint64), np.zeros(1, dtype=np.uint64))
    h, o = dbcache.load(d, [])
    assert h.shape == (0,) and o.tolist() == [0]
