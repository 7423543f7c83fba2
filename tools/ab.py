"""A/B of library variants on the GPU box: prints per-kernel ms of the reduced bench for each .so given.
Timing experiments may produce wrong numerics (max_sweeps is fixed so that the work is comparable)."""
import json, os, subprocess, sys
for lib in sys.argv[1:]:
    env = dict(os.environ, OAK_B200_LIB=os.path.abspath(lib), OAK_B200_FIXED_SWEEPS="8")
    p = subprocess.run([sys.executable, "bench.py", "--nx", "300", "--ny", "300", "--nobs", "90000", "--steps", "2", "--warmup", "1",
                        "--no-e2e", "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(lib, "value %.0f" % d["value"], d["roofline"]["kernel_ms_per_step"], "sweeps %.2f" % d["config"]["mean_jacobi_sweeps"], flush=True)
    except Exception as e:
        print(lib, "FAILED", e, p.stderr[-800:], flush=True)
