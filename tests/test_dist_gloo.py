"""Multi-process (world_size 2, gloo, CPU) test of the collective plumbing in symmer_b200.dist."""
import os
import socket
import subprocess
import sys

import pytest

from symmer_b200 import dist as sdist

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_block_bounds_and_log2():
    assert sdist.block_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert sdist.block_bounds(3, 8)[-1] == 3
    assert sdist.log2_exact(1) == 0 and sdist.log2_exact(8) == 3
    with pytest.raises(ValueError):
        sdist.log2_exact(6)


def test_owner_blocks():
    blocks = sdist.owner_blocks([2, 0, 3, 1], [1, 1, 1, 4], 2)
    # class a of A against class a ^ 2 of B
    assert blocks == [(0, 2, 2, 3), (2, 2, 3, 7), (2, 5, 0, 1), (5, 6, 1, 2)]
    total = sum(sum((p1 - p0) * (q1 - q0) for p0, p1, q0, q1 in sdist.owner_blocks([2, 0, 3, 1], [1, 1, 1, 4], r))
                for r in range(4))
    assert total == 6 * 7                                          # the owners tile the whole product


def test_single_process_passthrough():
    import torch
    t = torch.arange(6, dtype=torch.int64)
    assert sdist.exchange_records(t, torch.tensor([6])) is t
    full, offs = sdist.all_gather_rows(t.reshape(3, 2))
    assert offs == [0, 3] and full.shape == (3, 2)


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "_gloo_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=280)[0] for p in procs]
    for rank, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {rank} failed:\n{out}"
        assert f"rank {rank} ok" in out
