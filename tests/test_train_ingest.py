"""`yacht train` ingest (yacht_b200/utils.py: extract_signatures_and_info): unzip + gunzip + signature info in one
pass must leave the intermediate directory and the info dictionary exactly as the reference's three passes do
(make_training_data_from_sketches.py:109-119: extractall, gunzip pool, collect_signature_info)."""
import filecmp
import glob
import os
import zipfile

import numpy as np
import pytest

from yacht_b200 import sigio, synth, utils


def _three_passes(ref_zip, d, ksize, threads):
    with zipfile.ZipFile(ref_zip, "r") as z:
        z.extractall(d)
    utils.decompress_all_sig_files(glob.glob(f"{d}/signatures/*.sig.gz"), threads)
    return utils.collect_signature_info(threads, ksize, d)


def _tree(d):
    return sorted(os.path.relpath(os.path.join(r, f), d) for r, _, fs in os.walk(d) for f in fs)


@pytest.mark.parametrize("threads,n", [(1, 30), (4, 200)])
def test_one_pass_equals_three_passes(tmp_path, threads, n):
    db = synth.make_reference_db(n, 3, mean_size=150, sd_size=40)
    sk = [dict(name=f"genome {g}", mins=db.sketch(g), abundances=(np.arange(len(db.sketch(g))) % 7 + 1) if g % 3 == 0 else None)
          for g in range(db.n)]
    ref_zip = str(tmp_path / "db.zip")
    sigio.write_sig_zip(ref_zip, sk, 31)
    # members the reference also meets: a sketch of another k-mer size, an empty sketch, an uncompressed .sig, a stray file
    with zipfile.ZipFile(ref_zip, "a") as z:
        z.writestr("signatures/other_k.sig.gz", __import__("gzip").compress(sigio.signature_json("k21", [1, 2, 3], 21).encode()))
        z.writestr("signatures/empty.sig.gz", __import__("gzip").compress(sigio.signature_json("nothing", [], 31).encode()))
        z.writestr("signatures/plain.sig", sigio.signature_json("plain one", [5, 6, 7, 8], 31))
        z.writestr("signatures/README.txt", "not a signature")
        z.writestr("notes/extra.txt", "kept as is")
    a, b = str(tmp_path / "one"), str(tmp_path / "three")
    os.makedirs(a)
    os.makedirs(b)
    got = utils.extract_signatures_and_info(ref_zip, a, 31, threads)
    want = _three_passes(ref_zip, b, 31, threads)
    strip = lambda dd, root: {k: (v[0], v[1], v[2], v[3], os.path.relpath(v[4], root)) for k, v in dd.items()}
    assert strip(got, a) == strip(want, b)
    assert len(got) == n + 1 and "plain one" in got and "k21" not in got and "nothing" not in got
    assert _tree(a) == _tree(b)
    for rel in _tree(a):
        assert filecmp.cmp(os.path.join(a, rel), os.path.join(b, rel), shallow=False), rel
    assert not glob.glob(f"{a}/signatures/*.gz")                     # gunzipped in place, like the reference


def test_mean_abundance_and_scaled_fields(tmp_path):
    ref_zip = str(tmp_path / "db.zip")
    sigio.write_sig_zip(ref_zip, [dict(name="a", mins=[10, 20, 30], abundances=[1, 2, 6]), dict(name="b", mins=[11, 21], abundances=None)], 31)
    d = str(tmp_path / "x")
    os.makedirs(d)
    info = utils.extract_signatures_and_info(ref_zip, d, 31, 1)
    assert info["a"][1] == 3.0 and info["a"][2] == 3 and info["a"][3] == 1000 and info["b"][1] is None
    assert info["a"][0] == sigio.compute_md5sum(31, [10, 20, 30]) and os.path.basename(info["a"][4]) == info["a"][0] + ".sig"
