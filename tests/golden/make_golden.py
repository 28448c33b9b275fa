"""Generates tests/golden/ref_vlp16_seq.npz by running the REFERENCE'S OWN CODE (oracle/_ref: scanRegistration.cpp,
laserOdometry.cpp, laserMapping.cpp compiled unmodified from /root/reference) on the seeded synthetic sequence.

Only runs where /root/reference exists (this container).  The fixture travels; the reference does not.
    python tests/golden/make_golden.py
Inputs are not stored: the sweep generator is deterministic (gpscalibration_b200/csrc/synth.h, seed 0xC0FFEE), the
fixture stores a hash of every input sweep so a drifting generator is caught.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

N_SWEEPS = 24
HEAD = 48  # leading points of every feature cloud stored in clear


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from gpscalibration_b200 import SweepGenerator
    from oracle import ref
    assert ref.available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    gen = SweepGenerator(sensor=0, scene=0, seed=0xC0FFEE)
    ref.control_reset()
    names = ("full", "sharp", "less_sharp", "flat", "less_flat")
    out = {"n_sweeps": N_SWEEPS, "head": HEAD}
    counts = np.zeros((N_SWEEPS, 5), np.int32)
    odom = np.zeros((N_SWEEPS, 6), np.float32)
    rel = np.zeros((N_SWEEPS, 6), np.float32)
    mapped = np.zeros((N_SWEEPS, 6), np.float32)
    flags = np.zeros((N_SWEEPS, 3), np.int32)  # odom published, full-res published, mapping ran
    in_hash, cloud_hash, heads = [], [], []
    for k in range(N_SWEEPS):
        xyz = gen.sweep(k)[0].copy()
        in_hash.append(sha(xyz))
        r = ref.process(xyz, 100.0 + 0.1 * k)
        counts[k] = [f.shape[0] for f in r.features]
        cloud_hash.append([sha(f) for f in r.features])
        heads.append([f[:HEAD].copy() for f in r.features])
        odom[k], rel[k] = r.odom, r.rel
        flags[k] = [r.odom_published, r.fullres_published, r.mapping_ran]
        if r.mapping_ran:
            mapped[k] = r.mapped
        elif k:
            mapped[k] = mapped[k - 1]
    out.update(counts=counts, odom=odom, rel=rel, mapped=mapped, flags=flags, in_hash=np.array(in_hash),
               cloud_hash=np.array(cloud_hash), map_size=np.array(ref.map_size(), np.int32))
    for i, nm in enumerate(names):
        out["head_" + nm] = np.stack([np.pad(h[i], ((0, HEAD - h[i].shape[0]), (0, 0))) for h in heads])
    path = os.path.join(ROOT, "tests", "golden", "ref_vlp16_seq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; final odom", odom[-1], "final mapped", mapped[-1])


if __name__ == "__main__":
    main()
