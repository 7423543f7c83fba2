"""A/B of library variants on the GPU box: prints per-kernel ms of the reduced bench for each .so given.
Timing experiments may produce wrong numerics (max_sweeps is fixed so that the work is comparable)."""
import json, os, subprocess, sys
for lib in sys.argv[1:]:
    lib, _, zb = lib.partition(":")   # lib.so[:zones_per_batch]
    env = dict(os.environ, OAK_B200_LIB=os.path.abspath(lib), OAK_B200_FIXED_SWEEPS="8")
    if zb:
        env["OAK_B200_ZB"] = zb
    p = subprocess.run([sys.executable, "bench.py", "--nx", os.environ.get("AB_N", "300"), "--ny", os.environ.get("AB_N", "300"), "--nobs", str(int(os.environ.get("AB_N", "300")) ** 2), "--steps", "2", "--warmup", "1",
                        "--no-e2e", "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().splitlines()[-1])
        print(lib, zb, "value %.0f" % d["value"], d["roofline"]["kernel_ms_per_step"], "sweeps %.2f" % d["config"]["mean_jacobi_sweeps"], flush=True)
    except Exception as e:
        print(lib, "FAILED", e, p.stderr[-800:], flush=True)
