"""N1 measurement: the segment scheduler end to end on one GPU (and its two passes on two pipelines at once).

The reference's input_data node publishes at ros::Rate(IMRATE = 1 Hz) (IN:32, 268, 334) and waits for the SLAM nodes
between messages, so its replay rate is 1 sweep/s by construction; here the rate is whatever the pipeline sustains
through the blocking call plus one Python callback per message.
"""
import json, sys, time
sys.path.insert(0, '.')
from gpscalibration_b200 import SweepGenerator
from gpscalibration_b200.scheduler import SegmentScheduler

N = int(sys.argv[1]) if len(sys.argv) > 1 else 600
gen = SweepGenerator()
sweeps = [gen.sweep(k)[0].copy() for k in range(N)]
half = N // 2
bags = [sweeps[:half], sweeps[half:]]
stamps = [[10.0 + 0.1 * k for k in range(half)], [10.0 + 0.1 * k for k in range(half, N)]]
sched = SegmentScheduler(200.0, 80.0, 20.0)  # metres; the synthetic vehicle moves 1 m per sweep
out = {"sweeps_in_bags": N, "distances_m": [200.0, 80.0, 20.0]}
for mode in ("sequential", "parallel"):
    for rep in range(2):  # second pass = steady state (buffers grown)
        t0 = time.time()
        tracks, stats = sched.run(bags, stamps, parallel=(mode == "parallel"))
        dt = time.time() - t0
    pub = sum(s.published for s in stats)
    out[mode] = {"published_sweeps": int(pub), "resets": int(sum(s.resets for s in stats)), "tracks": len(tracks),
                 "seconds": round(dt, 3), "sweeps_per_s": round(pub / dt, 1)}
out["reference_rate_sweeps_per_s"] = 1.0
print(json.dumps(out))
