"""svk.data: the prefetcher and the asynchronous scalar reader move the same values in the same order."""
import pytest
import torch

import util  # noqa: F401  (sys.path set-up)
from svk.data import DevicePrefetcher, ScalarReader

pytestmark = pytest.mark.gpu


def test_device_prefetcher_preserves_batches_and_order():
    g = torch.Generator().manual_seed(0)
    batches = [(torch.randn(5, 7, 11, generator=g), torch.randint(0, 100, (5,), generator=g)) for _ in range(7)]
    seen = []
    for x, y in DevicePrefetcher(batches, "cuda"):
        assert x.is_cuda and y.is_cuda
        # consume on the compute stream with a kernel that takes a while, so the next copy really overlaps
        seen.append(((x * 1.0).cpu(), y.clone().cpu()))
    assert len(seen) == len(batches) and len(DevicePrefetcher(batches, "cuda")) == 7
    for (x, y), (xr, yr) in zip(seen, batches):
        assert torch.equal(x, xr) and torch.equal(y, yr)
    assert list(DevicePrefetcher([], "cuda")) == []


@pytest.mark.parametrize("n", [0, 1, 2, 3, 9])
def test_scalar_reader_returns_every_value_in_order(n):
    r = ScalarReader("cuda", depth=2)
    vals = [float(i) * 1.5 - 2.0 for i in range(n)]
    for v in vals:
        r.push(torch.tensor(v, device="cuda"))
    assert r.flush() == vals
    assert r.flush() == []
    r.push(torch.tensor(4.25, device="cuda"))       # usable again after a flush
    assert r.flush() == [4.25]
