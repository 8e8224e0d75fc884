"""Probe 3-D TMA tile loads: in-bounds, negative and beyond-the-edge coordinates, with and without 128B swizzle."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

def _probe_lib():
    """libclb_probe.so (clsurvey_b200/build.py build_probe): bring-up probes, not part of the product ABI."""
    import ctypes
    from clsurvey_b200 import build
    lib = ctypes.CDLL(build.build_probe() if not os.path.exists(os.path.join(build.OUT_DIR, "libclb_probe.so")) else os.path.join(build.OUT_DIR, "libclb_probe.so"))
    return lib


def _probe_call(name, *args):
    import ctypes
    fn = getattr(_probe_lib(), name)
    fn.restype = ctypes.c_int
    rc = fn(*[ctypes.c_void_p(a) if isinstance(a, int) and a > 2 ** 31 else a for a in args])
    assert rc == 0, (name, rc)
from clsurvey_b200 import _capi
_capi.lib()

def one(W, H, P, box, coords, swz):
    x = torch.arange(W * H * P, dtype=torch.float32).reshape(P, H, W).cuda() + 1.0
    out = torch.full((box[0] * box[1] * box[2],), -7.0, device="cuda")
    _probe_call("clb_debug_tma3d", x.data_ptr(), W, H, P, box[0], box[1], box[2], swz, coords[0], coords[1], coords[2],
               out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    o = out.cpu().reshape(box[2], box[1], box[0])
    # expected (no swizzle): zero outside
    ref = torch.zeros(box[2], box[1], box[0])
    for c in range(box[2]):
        for h in range(box[1]):
            for w in range(box[0]):
                pc, ph, pw = coords[2] + c, coords[1] + h, coords[0] + w
                if 0 <= pc < P and 0 <= ph < H and 0 <= pw < W:
                    ref[c, h, w] = x[pc, ph, pw].item()
    return o, ref

if __name__ == "__main__":
    case = sys.argv[1:]
    if case:
        W, H, P = 32, 32, 16
        coords = tuple(int(v) for v in case[0].split(","))
        swz = int(case[1])
        box = tuple(int(v) for v in case[2].split(",")) if len(case) > 2 else (32, 1, 8)
        o, ref = one(W, H, P, box, coords, swz)
        print("coords", coords, "swizzle", swz, "box", box, "match_unswizzled", bool(torch.equal(o, ref)),
              "same_multiset", bool(torch.equal(o.flatten().sort().values, ref.flatten().sort().values)))
        print(" row0", o[0, 0, :8].tolist(), " ref", ref[0, 0, :8].tolist())
    else:
        for args in (["0,0,0", "0"], ["0,0,0", "1"], ["-1,0,0", "0"], ["0,-1,0", "0"], ["1,0,0", "0"], ["-1,-1,0", "1"],
                     ["0,0,0", "1", "8,4,8"], ["-1,-1,0", "1", "8,4,8"], ["0,0,8", "1", "32,1,64"], ["-1,-1,0", "1", "32,1,64"]):
            r = subprocess.run([sys.executable, __file__] + args, capture_output=True, text=True, timeout=120)
            print((r.stdout.strip() or "") + ("" if r.returncode == 0 else "  [rc=%d] %s" % (r.returncode, r.stderr.strip().splitlines()[-1] if r.stderr.strip() else "")), flush=True)
