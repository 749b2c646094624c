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
