import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dpr_b200
from tests.helpers import make_inputs, rel_l2
from tests.gpu_util import dev_args, forced, to_dev, to_np
from oracle import oracle
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
dtype = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
grid = (64, 32, 32)
d = make_inputs(1, 3, 3, 3001, 2, grid, dtype, True)
args = dev_args(d, dtype)
td = torch.float32 if dtype == np.float32 else torch.float64
dpr_b200._lib.load().dpr_profile_enable(2)
with forced(forward_algo=3, pullback_algo=7):
    if which == "fwd":
        out = dpr_b200.raster(grid, *args); torch.cuda.synchronize()
        ref = oracle.raster(grid, d["points"], d["rotation"], d["translation"], d["background"], d["out_weight"], d["point_weight"], dtype=dtype, f64_accumulate=True)
        print("fwd ok", dpr_b200.last_path(0), rel_l2(to_np(out), ref))
    else:
        pb = dpr_b200.raster_pullback_(to_dev(d["ds_dout"], td), *args); torch.cuda.synchronize()
        print("pullback ok", dpr_b200.last_path(1))
