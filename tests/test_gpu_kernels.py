"""Per-kernel parity on the GPU, through the C ABI (see tests/kernel_checks.py for the individual checks)."""
import pytest

from tests import kernel_checks

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(kernel_checks.CHECKS))
def test_kernel(name):
    err, tol = kernel_checks.CHECKS[name]()
    assert err == err and err <= tol, f"{name}: error {err:.3e} exceeds tolerance {tol:.1e}"
