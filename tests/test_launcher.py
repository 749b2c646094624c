"""relightable_nr_b200.run: the reference's unchanged scripts import the drop-in modules.  The scripts are only
present in the build container (/root/reference); there every script is started through the launcher with --help, which executes
all of its top-level imports (network, render, neural_renderer, gcn_lib, ... + the compatibility shims) and argparse."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


def test_shims_and_registration():
    code = ("import sys; from relightable_nr_b200 import run, dropin; run.install_shims(); names = dropin.install();"
            "import network, render, camera, sph_harm, misc, neural_renderer, gcn_lib.dense, pytorch_prototyping.pytorch_prototyping;"
            "import tensorboardX, torch_geometric.data; import numpy as np; assert np.int is int;"
            "assert network.__name__.startswith('relightable_nr_b200.dropin'), network.__name__;"
            "d = torch_geometric.data.Data(pos=1, x=2); assert d.pos == 1;"
            "assert {'TextureMapper','RenderingNet','RaySampler','RayRenderer','LightingSH','LightingLP','Interpolater','Mesh'} <= set(dir(network));"
            "print(network.DenseDeepGCN.__name__, network.Rasterizer.__name__)")
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert 'DenseDeepGCN Rasterizer' in r.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference scripts are only present in the build container')
@pytest.mark.parametrize('script', ['train_rnr.py', 'test_rnr.py', 'train_dnr.py', 'test_dnr.py'])
def test_unchanged_script_starts_under_the_launcher(script):
    r = subprocess.run([sys.executable, '-m', 'relightable_nr_b200.run', os.path.join(REF, script), '--help'], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert 'usage: %s' % script in r.stdout and '--gpu_id' in r.stdout and 'drop-in modules registered' in r.stderr


TOY_SCRIPT = """
import os, torch
from torch.utils.data import DataLoader, TensorDataset
%s
loader = DataLoader(TensorDataset(torch.arange(8, dtype=torch.float32).view(8, 1)), batch_size=1, shuffle=True)
model = torch.nn.Linear(1, 3)
opt = torch.optim.Adam(model.parameters(), lr=1e-2)
seen = []
for (x,) in loader:
    seen.append(int(x.item())); opt.zero_grad(); (model(x) ** 2).sum().backward(); opt.step()
torch.save({'seen': seen, 'w': model.weight.detach().clone()}, os.path.join(r'%s', 'out_' + os.environ.get('RANK', '0') + '.pt'))
"""


@pytest.mark.parametrize('seeding', ['torch.manual_seed(0)', '', "torch.manual_seed(7 + int(os.environ.get('RANK', '0')))"],
                         ids=['seeded', 'unseeded', 'seeded-differently-per-rank'])
def test_launcher_data_parallel_mode(tmp_path, seeding):
    """torchrun + the launcher turn a plain single-process training script into a 2-rank data-parallel run (gloo on CPU):
    disjoint samples, replicas in sync.  The script below has no rank logic at all.  The reference scripts never seed
    (train_rnr.py has no torch.manual_seed), so the replicas must also end up identical when the script does not seed, or even
    seeds every rank differently: the launcher broadcasts rank 0's parameters when the optimizer is built."""
    script = tmp_path / 'toy_train.py'
    script.write_text(TOY_SCRIPT % (seeding, str(tmp_path)))
    import socket
    sock = socket.socket()
    sock.bind(('127.0.0.1', 0))
    port = sock.getsockname()[1]
    sock.close()
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', str(port), '-m', 'relightable_nr_b200.run', str(script)], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    import torch
    a = torch.load(str(tmp_path / 'out_0.pt'))
    b = torch.load(str(tmp_path / 'out_1.pt'))
    assert sorted(a['seen'] + b['seen']) == list(range(8)) and len(a['seen']) == 4
    assert torch.allclose(a['w'], b['w'], atol=1e-6)
