"""The `staged3` LoRA-merge pipeline (3-deep W ring decoupled from the 2-deep operand ring; the
default implementation since round 2) against the fp32 reference and, bit for bit, against the
`staged` pipeline.  Each case runs in a child process with a hard timeout: a tcgen05/mbarrier
pipeline that deadlocks costs one test, not the GPU session."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("CUDA device required", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_STAGED3 = r"""
import sys, torch
sys.path.insert(0, {root!r})
from rayfed_b200.ops import kernels as K
dev = "cuda:0"
P, r, m, n, has_base = {P}, {r}, {m}, {n}, {has_base}
torch.manual_seed(P * 100 + r)
Bs = [(torch.randn(m, r, device=dev) * 0.5).to(torch.bfloat16) for _ in range(P)]
Ats = [(torch.randn(n, r, device=dev) * 0.5).to(torch.bfloat16) for _ in range(P)]
W = torch.randn(m, n, device=dev).to(torch.bfloat16)
dw = torch.cat(Bs, 1).float() @ torch.cat(Ats, 1).float().t()
ref = ((W.float() if has_base else 0) + dw / P).to(torch.bfloat16)
for rep in range(2):
    out = W.clone()
    K.lora_merge(b_ptrs=[b.data_ptr() for b in Bs], at_ptrs=[a.data_ptr() for a in Ats], out_ptr=out.data_ptr(),
                 m=m, n=n, r=r, w_base_ptr=out.data_ptr() if has_base else 0, impl="staged3")
    torch.cuda.synchronize()
    err = (out.float() - ref.float()).abs().max().item()
    assert err <= 2 ** -6 * max(1.0, ref.float().abs().max().item()), err
    assert (out.float() - ref.float()).abs().mean().item() < 1e-3
# and it must agree with the validated pipeline bit for bit
a = W.clone(); b = W.clone()
for impl, o in (("staged", a), ("staged3", b)):
    K.lora_merge(b_ptrs=[x.data_ptr() for x in Bs], at_ptrs=[x.data_ptr() for x in Ats], out_ptr=o.data_ptr(),
                 m=m, n=n, r=r, w_base_ptr=o.data_ptr() if has_base else 0, impl=impl)
torch.cuda.synchronize()
assert torch.equal(a, b)
print("ok")
"""


@pytest.mark.parametrize("P,r,m,n,has_base", [
    (2, 16, 256, 256, True),      # one tile per CTA, K = 32 (one atom)
    (8, 16, 4096, 2048, True),    # 512 tiles: every ring wraps (W: 3 stages, operands: 2, TMEM: 2)
    (8, 16, 4096, 2048, False),   # no W load: the stage hand-over is a plain arrive
    (4, 32, 200, 192, True),      # partial tiles, TMA clipping
    (1, 64, 384, 96, True),
])
def test_lora_merge_staged3_matches_reference_and_staged(P, r, m, n, has_base):
    code = _STAGED3.format(root=ROOT, P=P, r=r, m=m, n=n, has_base=has_base)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0 and "ok" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
