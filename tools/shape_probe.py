import os, sys
sys.path.insert(0, "/root/repo")
from nbmf_b200 import Engine, _lib
for (F, N) in [(262144, 32768), (262144, 65536)]:
    e = Engine(precision=_lib.PREC_TENSOR_FAST)
    e.generate_synthetic(F, N, 8, 1.0, 1.0, 5); e.init_random_labels(128, 6)
    e.vb_begin(128, 1.0, 1.0, 1.0/128)
    for _ in range(6): e.vb_step(False)
    e.vb_end(0, want_outputs=False)
    t = e.timing(); e.close()
    print(F, N, "rows %.2f cols %.2f params %.2f" % (t["pass_rows_ms"], t["pass_cols_ms"], t["params_ms"]), flush=True)
