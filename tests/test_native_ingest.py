"""The native signature scanner (yacht_b200/csrc/sig_scan.hpp, behind ygpu_read_signatures) on hand-made documents:
what the reference reads is document[0]["signatures"][0]["mins"] through nlohmann::json (src/cpp/main.cpp:62-84).
The number path converts digits eight at a time; every literal length, position and separator style must give
the value Python's own parser gives.  No GPU needed: ingest is host code."""
import json
import os

import numpy as np
import pytest

from yacht_b200 import _lib


def _doc(mins_text: str, extra_before: str = "", extra_after: str = "") -> str:
    return ('[{"class":"sourmash_signature","name":"x","signatures":[{' + extra_before + '"ksize":31,"mins":' + mins_text
            + extra_after + ',"md5sum":"0"}],"version":0.4}]')


def _read(tmp_path, texts, threads=2):
    paths = []
    for k, t in enumerate(texts):
        p = str(tmp_path / f"s{k}.sig")
        with open(p, "w") as f:
            f.write(t)
        paths.append(p)
    h, o, bad = _lib.read_signatures(paths, threads)
    return [h[int(o[k]):int(o[k + 1])].tolist() for k in range(len(texts))], bad


def test_every_literal_length_and_separator_style(tmp_path):
    rng = np.random.default_rng(0)
    vals = []
    for digits in range(1, 21):                                   # 1 .. 20 decimal digits
        lo, hi = 10 ** (digits - 1), min(10 ** digits, 2 ** 64) - 1
        vals += [lo if digits > 1 else 0, hi] + [int(rng.integers(lo, hi, dtype=np.uint64)) for _ in range(5)]
    vals += [2 ** 64 - 1, 18446744073709552, 0, 7, 99999999, 100000000, 9999999999999999, 10000000000000000]
    compact = "[" + ",".join(str(v) for v in vals) + "]"
    spaced = "[ " + " , ".join(str(v) for v in vals) + " ]"
    lines = "[\n" + ",\n".join("  " + str(v) for v in vals) + "\n]"
    got, bad = _read(tmp_path, [_doc(compact), _doc(spaced), _doc(lines), _doc(compact, extra_before='"abundances":[1,2,3],"note":"mins\\":[9]",')])
    assert bad == 0
    for g in got:
        assert g == vals


def test_short_documents_and_array_edges(tmp_path):
    # literals within 16 bytes of the end of the buffer take the scalar tail; empty arrays; a single value
    texts = ['[{"signatures":[{"mins":[1]}]}]', '[{"signatures":[{"mins":[]}]}]', '[{"signatures":[{"mins":[ ]}]}]',
             '[{"signatures":[{"mins":[12345678]}]}]', '[{"signatures":[{"mins":[123456789012345678,5]}]}]',
             '[{"signatures":[{"mins":[12345678,87654321,1]}]}]']
    got, bad = _read(tmp_path, texts, threads=1)
    assert got == [[1], [], [], [12345678], [123456789012345678, 5], [12345678, 87654321, 1]] and bad == 0


def test_non_plain_literals_follow_the_reference_cast(tmp_path):
    # nlohmann hands back whatever number it finds, cast to hash_t (main.cpp:78-81): floats truncate, negatives wrap
    got, _ = _read(tmp_path, [_doc("[1.5e3,12.0,-1,3]")])
    assert got == [[1500, 12, 2 ** 64 - 1, 3]]


def test_only_first_record_and_first_subsignature(tmp_path):
    text = ('[{"signatures":[{"ksize":21,"mins":[5,6]},{"ksize":31,"mins":[7]}]},{"signatures":[{"mins":[8]}]}]')
    got, _ = _read(tmp_path, [text])
    assert got == [[5, 6]]


def test_unreadable_and_malformed(tmp_path):
    good = str(tmp_path / "g.sig")
    with open(good, "w") as f:
        f.write(_doc("[3,4]"))
    h, o, bad = _lib.read_signatures([good, str(tmp_path / "missing.sig")], 2)
    assert bad == 1 and h.tolist() == [3, 4] and o.tolist() == [0, 2, 2]          # missing file = empty sketch (main.cpp:68-71)
    broken = str(tmp_path / "b.sig")
    with open(broken, "w") as f:
        f.write('[{"name":"x"}]')
    with pytest.raises(_lib.YgpuError):
        _lib.read_signatures([broken], 1)


def test_agrees_with_json_module_on_random_documents(tmp_path):
    rng = np.random.default_rng(5)
    texts, want = [], []
    for k in range(40):
        n = int(rng.integers(0, 300))
        bits = int(rng.integers(1, 65))
        v = [int(x) for x in rng.integers(0, 2 ** bits - 1, size=n, dtype=np.uint64, endpoint=True)]
        sep = [",", ", ", " ,\n"][k % 3]
        texts.append(_doc("[" + sep.join(map(str, v)) + "]"))
        want.append(json.loads(texts[-1])[0]["signatures"][0]["mins"])
    got, _ = _read(tmp_path, texts, threads=3)
    assert got == want


def test_ksize_selecting_reader(tmp_path):
    """ygpu_read_signatures_ksize: the run side's rule (reference utils.py:31-51) -- all records and sub-signatures are
    candidates, exactly one must have the requested k-mer size."""
    import json
    from yacht_b200 import _lib
    def sub(k, mins):
        return {"num": 0, "ksize": k, "seed": 42, "max_hash": 18446744073709552, "mins": mins, "md5sum": "x", "molecule": "dna"}
    a = tmp_path / "a.sig"
    a.write_text(json.dumps([{"class": "sourmash_signature", "name": "a", "signatures": [sub(21, [1, 2, 3]), sub(31, [10, 20, 30, 40])], "version": 0.4}]))
    b = tmp_path / "b.sig"     # "mins" before "ksize", second record holds the match
    b.write_text(json.dumps([{"name": "b0", "signatures": [{"mins": [7], "ksize": 51}]},
                             {"name": "b1", "signatures": [{"mins": [5, 6], "abundances": [1, 1], "ksize": 31}]}]))
    h, off, bad = _lib.read_signatures([str(a), str(b)], 2, ksize=31)
    assert bad == 0 and off.tolist() == [0, 4, 6] and h.tolist() == [10, 20, 30, 40, 5, 6]
    h, off, bad = _lib.read_signatures([str(a), str(b)], 2)              # the train core's rule: first record, first sub-signature
    assert off.tolist() == [0, 3, 4] and h.tolist() == [1, 2, 3, 7]
    with pytest.raises(_lib.YgpuError, match="Expected exactly one signature with ksize 41, found 0"):
        _lib.read_signatures([str(a)], 1, ksize=41)
    c = tmp_path / "c.sig"
    c.write_text(json.dumps([{"name": "c", "signatures": [sub(31, [1]), sub(31, [2])]}]))
    with pytest.raises(_lib.YgpuError, match="found 2"):
        _lib.read_signatures([str(c)], 1, ksize=31)
