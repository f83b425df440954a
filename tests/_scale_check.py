"""Scale check (run manually on the GPU box): the drop-in executable against the UNMODIFIED reference core on the
same N-genome synthetic signature files -- outputs compared byte for byte, wall times side by side.
usage: python tests/_scale_check.py [N] [threads]"""
import os, sys, time, glob, shutil, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yacht_b200 import synth
from oracle import train_oracle as to

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
thr = 0.95 ** 31
root = tempfile.mkdtemp(prefix="yacht_scale_")
t0 = time.time(); db = synth.make_reference_db(n, 2); print(f"generated {n} genomes, {int(db.offsets[-1])} hashes in {time.time()-t0:.1f}s", flush=True)
t0 = time.time(); paths = to.write_sig_dir(db.hashes, db.offsets, root); print(f"wrote {n} .sig files in {time.time()-t0:.1f}s", flush=True)
fl = os.path.join(root, "training_sig_files.tsv")
exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "yacht_b200", "run_yacht_train_core")
res = {}
for name, binary in (("b200", exe), ("reference", to.REF_BIN)):
    wd = os.path.join(root, "wd_" + name); os.makedirs(wd)
    sel = os.path.join(wd, "selected_result.tsv")
    t0 = time.time()
    cp = subprocess.run([binary, "-t", str(threads), "-c", repr(thr), "-p", "1", fl, wd, sel], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    assert cp.returncode == 0, cp.stderr[-2000:]
    lines = []
    for f in sorted(glob.glob(os.path.join(wd, "*_*.txt"))):
        lines.extend(open(f).read().splitlines())
    res[name] = dict(wall=wall, lines=sorted(lines), selected=open(sel).read(), phases=to.parse_phase_times(cp.stdout),
                     tail=[l for l in cp.stdout.splitlines() if l.startswith("[gpu")])
    print(name, f"wall {wall:.2f}s", res[name]["phases"], res[name]["tail"], flush=True)
same_pairs = res["b200"]["lines"] == res["reference"]["lines"]
same_sel = res["b200"]["selected"] == res["reference"]["selected"]
print(f"N={n} threads={threads}: pair lines identical={same_pairs} ({len(res['reference']['lines'])} lines), "
      f"selected_result.tsv identical={same_sel} ({res['reference']['selected'].count(chr(10))} genomes kept), "
      f"speed-up (whole executable, files to files) = {res['reference']['wall'] / res['b200']['wall']:.1f}x")
shutil.rmtree(root, ignore_errors=True)
assert same_pairs and same_sel
