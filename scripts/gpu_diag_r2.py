"""Round-2 diagnosis of the six GPU failures of round 1 (VERDICT "What's weak" 1-2): dumps what the CUDA build computes so that
the comparison with the oracles / host mirror can be done offline.  Writes gpurun_out/r2_diag_*.npz and prints tracebacks."""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import numpy as np, torch
from common import load_golden, native_case, rel_err
os.makedirs(ROOT + "/gpurun_out", exist_ok=True)

# ---- 1. FP32FP16 + Regularized / Outflow / Halfway: every variant, every step ------------------------------------------
out = {}
for name in ("sphere_d3q19_bgk_fp32fp16",):
    g = load_golden(name)
    for backend, v in (("WARP", 1), ("WARP", 202), ("WARP", 203), ("WARP", 2), ("JAX", 0)):
        try:
            stepper, f_0, f_1, bm, mm = native_case(g, backend=backend, cells_per_thread=v)
            hist = [f_0.numpy().copy()]
            aux = [f_1.numpy()[0].copy()]
            for i in range(g["steps"]):
                f_0, f_1 = stepper(f_0, f_1, bm, mm, g["omega"], i)
                f_0, f_1 = f_1, f_0
                hist.append(f_0.numpy().copy())
            out[f"{name}|{backend}|{v}"] = np.stack(hist)
            out[f"{name}|{backend}|{v}|aux0"] = aux[0]
            print(name, backend, v, "rel err vs golden %.3e" % rel_err(hist[-1], g["f_final"]), flush=True)
        except Exception:
            traceback.print_exc()
np.savez_compressed(ROOT + "/gpurun_out/r2_diag_fp16.npz", **out)

# ---- 2. stand-alone extended collision operators: full tracebacks + the arrays ----------------------------------------
import xlb_b200 as xlb
from xlb_b200.compute_backend import ComputeBackend
from xlb_b200.operator.collision import BGK, KBC, ForcedCollision, SmagorinskyLESBGK
from xlb_b200.operator.force import ExactDifference
from oracle import lbm_numpy as O
pp, be = xlb.PrecisionPolicy.FP32FP32, ComputeBackend.WARP
out = {}
for lattice in ("D3Q19", "D3Q27", "D2Q9"):
    try:
        vs = getattr(xlb.velocity_set, lattice)(pp, be)
        xlb.init(velocity_set=vs, default_backend=be, default_precision_policy=pp)
        lat = O.Lattice(lattice)
        shape = (6, 5, 4) if lat.d == 3 else (9, 7)
        rng = np.random.default_rng(3)
        rho = (1.0 + 0.01 * rng.standard_normal((1,) + shape)).astype(np.float32)
        u = (0.03 * rng.standard_normal((lat.d,) + shape)).astype(np.float32)
        feq = O.equilibrium(rho, u, lat)
        f = (feq * (1.0 + 0.02 * rng.standard_normal(feq.shape))).astype(np.float32)
        force = np.array([2e-4, -1e-4, 5e-5][: lat.d])
        dev = lambda a: torch.as_tensor(a if lat.d == 3 else a[..., None]).cuda()
        back = lambda t: t.cpu().numpy() if lat.d == 3 else t.cpu().numpy()[..., 0]
        F, FEQ, RHO, U = dev(f), dev(feq), dev(rho), dev(u)
        out[f"{lattice}|f"], out[f"{lattice}|feq"], out[f"{lattice}|rho"], out[f"{lattice}|u"] = f, feq, rho, u
        def rec(key, fn, want):
            try:
                got = back(fn())
                out[f"{lattice}|{key}|got"], out[f"{lattice}|{key}|want"] = got, want
                print(lattice, key, "full rel err %.3e  increment rel err %.3e" % (rel_err(got, want), rel_err(got - f, want - f)), flush=True)
            except Exception:
                print(lattice, key, "ERROR", flush=True)
                traceback.print_exc()
        rec("ed", lambda: ExactDifference(force)(F, FEQ, torch.empty_like(F), RHO, U), O.exact_difference_force(f.copy(), feq, rho, u, force, lat))
        cases = [("BGK", BGK)] + ([("KBC", KBC)] if lattice != "D3Q19" else []) + ([("SmagorinskyLESBGK", SmagorinskyLESBGK)] if lat.d == 3 else [])
        for cname, cls in cases:
            if cname == "BGK":
                base = O.collide_bgk(f, feq, 1.7)
            elif cname == "KBC":
                base = O.collide_kbc(f, feq, rho, lat, 1.7)
            else:
                base = O.collide_smagorinsky(f, feq, lat, 1.7)
                rec(cname, lambda: cls()(F, FEQ, RHO, U, torch.empty_like(F), 1.7), base)
            want = O.exact_difference_force(base, feq, rho, u, force, lat)
            rec("forced_" + cname, lambda: ForcedCollision(cls(), force_vector=force)(F, FEQ, torch.empty_like(F), RHO, U, 1.7), want)
    except Exception:
        print(lattice, "SETUP ERROR", flush=True)
        traceback.print_exc()
np.savez_compressed(ROOT + "/gpurun_out/r2_diag_ops.npz", **out)
