"""Run every kernel parity check and print a table (does not stop at the first failure).  GPU box only."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from tests import kernel_checks  # noqa: E402

if __name__ == "__main__":
    only = sys.argv[1:]
    if only:
        for k in [k for k in kernel_checks.CHECKS if not any(o in k for o in only)]:
            del kernel_checks.CHECKS[k]
    print(torch.cuda.get_device_name(0))
    res = kernel_checks.run_all()
    bad = [r for r in res if r[3] != "ok"]
    print(f"{len(res) - len(bad)}/{len(res)} ok")
    sys.exit(1 if bad else 0)
