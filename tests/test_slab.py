"""Slab decomposition over several ranks (one process per GPU; replaces the reference's Universe / Domain sub-boxes,
Grid::reduce_ghost_nodes src/grid.cpp:477-621,881-1132 and ULMPM::exchange_particles src/ulmpm.cpp:565-667).

CPU (gloo, world_size 2 and 3): the host-side partition.  GPU (NCCL, >= 2 devices): stepping parity with the
single-rank oracle, with particles migrating across the slab cuts.
"""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = os.path.join(ROOT, "tests", "slab_worker.py")


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def launch(world, args, timeout=600):
    for attempt in range(3):  # the probed port can be taken again before the rendezvous binds it
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
               "--master-port", str(free_port()), WORKER] + args
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
        if not (p.returncode != 0 and "EADDRINUSE" in p.stderr):
            break
    if not (p.returncode == 0 and "SLAB-OK" in p.stdout):
        fails = [ln for ln in p.stdout.splitlines() if ln.startswith("SLAB-FAIL")]
        raise AssertionError("worker failed (rc %d)\n%s\n%s" % (p.returncode, "\n".join(fails) or p.stdout[-1500:], p.stderr[-1500:]))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):  # keep the measured errors of passing runs (pytest -q hides them)
        with open(os.path.join(out_dir, "slab_results.log"), "a") as f:
            f.write("".join("%s | %s\n" % (" ".join(args), ln) for ln in p.stdout.splitlines() if ln.startswith("SLAB-OK")))
    return p.stdout


@pytest.mark.parametrize("world,shape", [(2, "cubic-spline"), (3, "cubic-spline"), (2, "linear")])
def test_partition_cpu(oracle_lib, world, shape):
    launch(world, ["partition", "--cells", "12", "5", "4", "--shape", shape])


def test_partition_second_solid_cpu(oracle_lib):
    """Several solids on one decomposed grid: the second solid is split by the first solid's cuts, a slab may hold none of its particles, tags
    continue after the first solid's global count."""
    out = launch(3, ["partition", "--cells", "24", "5", "4", "--variant", "two_solids"])
    assert "SLAB-OK second solid" in out


def test_thin_slab_is_rejected(oracle_lib):
    """A slab thinner than the stencil overlap cannot own its shared planes: the host driver refuses the run."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "4", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), WORKER, "partition", "--cells", "4", "3", "3"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0 and "thinner than the stencil" in (p.stdout + p.stderr)


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.gpu
@pytest.mark.parametrize("scheme,drift", [("musl", 0.03), ("usl", 0.0)])
def test_two_slabs_match_single_rank_oracle(oracle_lib, scheme, drift):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = launch(2, ["step", "--cells", "12", "6", "6", "--scheme", scheme, "--drift", str(drift), "--steps", "100", "--a", "2.5e-3"])
    print(out)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["APIC, cubic-spline", "MLS, quadratic-spline", "FLIP, cubic-spline, 0.99, mechanical, gradient-enhanced"])
def test_two_slabs_affine_transfer(oracle_lib, method):
    """APIC family / gradient-enhanced projection over two slabs: the stored velocity gradient migrates with the particle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    print(launch(2, ["step", "--cells", "12", "6", "6", "--scheme", "musl", "--drift", "0.03", "--steps", "100", "--a", "2.5e-3", "--method", method]))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["velocity_nodes", "thermal", "force_nodes", "delete_particles"])
def test_two_slabs_node_fix_and_thermal_fields(oracle_lib, variant):
    """fix velocity_nodes on a node group that spans both slabs (reaction force all-reduced) and a thermo-mechanical block (T, Qext,
    Qint in the halo sums), with particles migrating."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    print(launch(2, ["step", "--cells", "12", "6", "6", "--scheme", "musl", "--drift", "0.03", "--steps", "100", "--a", "2.5e-3", "--variant", variant]))


@pytest.mark.gpu
@pytest.mark.parametrize("world,cells", [(2, "12"), (4, "24")])
@pytest.mark.parametrize("variant", ["two_solids", "rigid_tool"])
def test_slabs_second_solid(oracle_lib, world, cells, variant):
    """A second solid on the same decomposed grid, compared solid by solid: a deformable plate pressed onto the drifting block, and the same
    plate made rigid (Grid::rigid flags OR-ed over the shared planes, Grid::reduce_rigid_ghost_nodes src/grid.cpp:746-879).  On 4 slabs the
    outer two hold no particle of the plate."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    print(launch(world, ["step", "--cells", cells, "6", "6", "--scheme", "musl", "--drift", "0.03", "--steps", "100", "--a", "2.5e-3", "--variant", variant]))


@pytest.mark.gpu
@pytest.mark.parametrize("a", ["2.5e-4", "2.5e-3"])
def test_four_slabs_match_single_rank_oracle(oracle_lib, a):
    """2.5e-4 is the benchmark's own squeeze rate; 2.5e-3 is ten times that (round 1 measured the stress of this block at
    1.03e-10 from the oracle at that rate and lowered the rate - it is back)."""
    if _ngpu() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    print(launch(4, ["step", "--cells", "24", "6", "6", "--scheme", "musl", "--drift", "0.03", "--steps", "100", "--a", a]))


@pytest.mark.gpu
def test_eight_slabs_match_single_rank_oracle(oracle_lib):
    """the configuration the scaling run times: 8 slabs, particles drifting across every cut"""
    if _ngpu() < 8:
        pytest.skip("needs 8 GPUs (gpurun --gpus 8)")
    print(launch(8, ["step", "--cells", "48", "6", "6", "--scheme", "musl", "--drift", "0.03", "--steps", "100", "--a", "2.5e-4"]))
