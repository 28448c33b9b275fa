"""Generates tests/golden/ref_imu_seq.npz by running the REFERENCE'S OWN CODE (private copies of oracle/_ref/libref_sr.so and
libref_lo.so: scanRegistration.cpp with its IMU branch, laserOdometry.cpp fed with the /imu_trans it publishes) on the seeded
IMU scenarios of tests/test_imu_deskew.py.

Only runs where /root/reference exists (this container).  The fixture travels; the reference does not.
    python tests/golden/make_golden_imu.py
Stored per scenario and sweep: the twelve floats of /imu_trans, sizes + SHA-256 of the five feature clouds, the odometry
node's transformSum / transformation / publish flags and the SHA-256 of the clouds it published; plus a hash of every input
(sweep and IMU message) so a drifting generator is caught.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEEDS = (0, 1)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import ref
    from test_imu_deskew import scenario
    assert ref.available(), "build oracle/_ref first (python -c 'import __graft_entry__ as g; g.build()')"
    out = {"seeds": np.array(SEEDS, np.int32)}
    for seed in SEEDS:
        sr, lo = ref.SrWithImu(), ref.LoWithImu()
        tr_all, counts, chash, lo_out, lo_hash, in_hash = [], [], [], [], [], hashlib.sha256()
        for ev in scenario(seed):
            if ev[0] == "imu":
                for v in ev[1:]:
                    in_hash.update(np.ascontiguousarray(v, np.float64).tobytes())
                sr.imu(*ev[1:])
                continue
            _, stamp, xyz = ev
            in_hash.update(np.ascontiguousarray(xyz, np.float32).tobytes())
            feat, tr = sr.process(xyz, stamp)
            o, clouds = lo.step(feat, stamp, tr)
            tr_all.append(tr)
            counts.append([f.shape[0] for f in feat])
            chash.append([sha(f) for f in feat])
            lo_out.append(o[:15])
            lo_hash.append([sha(c) for c in clouds] if clouds is not None else ["", "", ""])
        lo.close()
        p = f"s{seed}_"
        out.update({p + "imu_trans": np.array(tr_all, np.float32), p + "counts": np.array(counts, np.int32), p + "cloud_hash": np.array(chash),
                    p + "lo_out": np.array(lo_out, np.float32), p + "lo_cloud_hash": np.array(lo_hash), p + "in_hash": np.array(in_hash.hexdigest())})
    path = os.path.join(ROOT, "tests", "golden", "ref_imu_seq.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
