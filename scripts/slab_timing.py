"""Per-role cycle breakdown of the slab aggregation kernel (gcs_debug_slab_timing)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gcn_string_b200 as g
from gcn_string_b200 import _lib, ops
from spmm_bench import make_batch

lib = _lib.load()
lib.gcs_debug_slab_timing.argtypes = [ctypes.c_void_p]
lib.gcs_debug_slab_timing.restype = None
H = 256
a = make_batch(1024, 500, 12)
x = torch.randn(a.n_rows, H, device="cuda")
y = torch.empty(a.n_rows, 1280, device="cuda")[:, :H]
sc, sh, al = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda"), torch.rand(H, device="cuda") * 0.3
for stages in (2,):
    lib.gcs_debug_set_param(10, stages)
    for hgt in (4, 2):
        for tr in (True, False):
            args = (sc, sh, al) if tr else (None, None, None)
            rb = a.rb(hgt)
            call = lambda: ops.spmm_sum_graphs(a.graph_ptr, a.max_graph_nodes, a.rowptr, a.colidx, x, *args, out=y, rb=rb, rb_height=hgt)
            call(); torch.cuda.synchronize()
            cnt = torch.zeros(8, dtype=torch.int64, device="cuda")
            lib.gcs_debug_slab_timing(cnt.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            lib.gcs_debug_slab_timing(None)
            c = cnt.cpu().numpy().astype(np.float64)
            items = c[6]
            n_cta = 148
            out = {"stages": stages, "rb": hgt, "prologue": tr, "us": round(e0.elapsed_time(e1) * 1e3, 1), "items": int(items),
                   "producer_wait_cyc_per_item": round(c[0] / items), "producer_work_cyc_per_item": round(c[1] / items),
                   "xform_wait_cyc_per_item_per_warp": round(c[2] / items / 8), "xform_work_cyc_per_item_per_warp": round(c[3] / items / 8),
                   "gather_wait_cyc_per_item_per_warp": round(c[4] / items / 23), "gather_work_cyc_per_item_per_warp": round(c[5] / items / 23)}
            print(json.dumps(out), flush=True)
