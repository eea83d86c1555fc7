"""GPU suite, multi-GPU part (SURVEY.md section 4 bullet 4 / section 8e): a book sharded over G GPUs with NCCL
scatter / gather (tools/shard_run.py under torchrun) must give byte-identical mask / fg / bg for every world size.
Needs at least 2 GPUs (skipped on a single-GPU box); config-3 pages (dpi 300, window 75, 30 % halftone pages)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


def _run(world, pages, shape=None):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()),
           os.path.join(ROOT, 'tools', 'shard_run.py'), '--pages', str(pages), '--config', '3', '--reps', '1']
    if shape:
        cmd += ['--shape', str(shape[0]), str(shape[1])]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert lines, out.stdout[-2000:]
    return json.loads(lines[-1])


def test_sharded_book_is_byte_identical_across_world_sizes():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    worlds = [w for w in (1, 2, 4, 8) if w <= n]
    full = os.environ.get('B200MRC_SHARD_FULL_SIZE', '1') == '1'
    shape = None if full else (660, 510)                      # config 3 at full size: 3300 x 2550
    res = {w: _run(w, 24, shape) for w in worlds}
    ref = res[1]['sha256']
    assert set(ref) == {'mask', 'fg', 'bg'}
    for w in worlds[1:]:
        assert res[w]['sha256'] == ref, (w, res[w]['sha256'], ref)
        assert res[w]['pre_sharded_Mpx_s'] > 0 and res[w]['root_scatter_included_Mpx_s'] > 0
