"""Small driver for ncu: one config-2-sized call (the kernels of the path at the headline shape), so that
`ncu --set full -k regex:...` can capture the CUDA-core kernels' DRAM traffic.  Run under gpurun; see tools/run_ncu.sh."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import speech_resynth_b200 as srb  # noqa: E402
from speech_resynth_b200 import synthetic  # noqa: E402

if __name__ == "__main__":
    nfe = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    decoder = srb.ConditionalFlowMatchingWithHifiGan(srb.reference_config()).eval()
    decoder.load_state_dict(synthetic.make_state_dict(0), strict=True)
    decoder = decoder.cuda()
    os.environ.setdefault("SRB_GRAPHS", "0")
    ids = synthetic.make_units(64, 500, seed=7)
    for _ in range(2):
        decoder(ids, 1.0 / nfe, 1.0)
    torch.cuda.synchronize()
