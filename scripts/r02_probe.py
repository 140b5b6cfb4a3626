import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import numpy as np, torch, scipy.sparse as sp
    from gcn_string_b200 import _lib, ops
    S, sb, n, H, variant = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    lib = _lib.load()
    lib.gcs_debug_set_param(10, S); lib.gcs_debug_set_param(11, sb)
    rng = np.random.default_rng(0)
    sizes = [n, n + 1]
    mats = [sp.csr_matrix((rng.random((m, m)) < 0.1).astype(np.int64)) for m in sizes]
    a = sp.block_diag(mats, format="csr"); a.sort_indices()
    gp = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)).cuda()
    rp, ci = torch.tensor(a.indptr.astype(np.int32)).cuda(), torch.tensor(a.indices.astype(np.int32)).cuda()
    x = torch.randn(a.shape[0], H, device="cuda")
    rb = ops.build_rb(rp, ci, 4)
    torch.cuda.synchronize()
    sc, sh, al = torch.rand(H, device="cuda") + 0.5, torch.randn(H, device="cuda"), torch.rand(H, device="cuda")
    tr = (sc, sh, al) if "t" in variant else (None, None, None)
    res = torch.randn(a.shape[0], H, device="cuda") if "r" in variant else None
    out = torch.zeros(a.shape[0], 3 * H, device="cuda")[:, H:2 * H] if "o" in variant else None
    y = ops.spmm_sum_graphs(gp, max(sizes), rp, ci, x, *tr, residual=res, out=out, rb=rb, rb_height=4)
    torch.cuda.synchronize()
    lib.gcs_debug_set_spmm_mode(1)
    ref = ops.spmm_sum(rp, ci, x, *tr)
    if res is not None: ref = ref + res
    print("OK", bool(torch.equal(ref, y)), int(lib.gcs_spmm_slab_stage_bytes()))
else:
    for S in (2,):
        for sb in (32768, 115584):
            for n, H in ((150, 32), (150, 512)):
                for variant in ("-", "t", "r", "o", "tro"):
                    r = subprocess.run([sys.executable, __file__, str(S), str(sb), str(n), str(H), variant], capture_output=True, text=True)
                    out = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip().splitlines()[-1][:120]
                    print(S, sb, n, H, variant, "->", out, flush=True)
