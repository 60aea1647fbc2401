"""compute-sanitizer target: the kernels added in the second session of round 2 on small grids — odd-length family (255 on
each axis, odd row count), three-pass last-axis kernels at 512 / 1024 with the new line layout, two-stage S3 at 512,
fh_download.  Checks the operator against the oracle so a silent mis-address would also show as a wrong result."""
import sys, os
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'oracle')); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import ffthom_oracle as O, harness
from ffthompy_b200 import device
from ffthompy_b200.tensors import Tensor
device.init(0)
for N in [(3, 5, 255), (255, 2, 8), (2, 255, 8), (2, 4, 512), (2, 4, 1024), (512, 2, 8)]:
    rng = np.random.default_rng(sum(N))
    G = harness.green_for('scalar', 'GaNi', N, np.ones(3), 'primal')[0]
    Go = O.proj_scalar(N, np.ones(3))[1]
    Aval = np.einsum('ij,...->ij...', np.eye(3), 1.+10.*(rng.random(N) < 0.3))
    A, Afun = harness.build_operator(Aval, G, np.array(N))
    u = rng.standard_normal((3,)+N)
    got = Afun(Tensor(name='u', val=u, order=1, N=np.array(N))).val
    ref = O.GA(Aval, Go, N)(u)
    err = np.abs(got-ref).max()/np.abs(ref).max()
    print(N, Afun.fused().config()['last'], Afun.fused().config()['mid0'], 'err %.2e' % err)
    assert err < 1e-12
import torch
t = torch.randn((3, 700001), dtype=torch.float64, device=device.device())
assert np.array_equal(device.download(t), t.cpu().numpy())
print('SANITIZE TARGET OK')
