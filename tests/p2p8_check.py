"""Minimal many-rank check of the peer-memory mailboxes (csrc/comm.cu): a decomposed c1_64k LJ liquid with the
Bussi thermostat (both one-word reductions every step) runs 80 steps; the mailboxes must be active and the
kinetic energy finite and equal on all ranks. Launched under torchrun with one rank per GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import hoomd_blue_b200 as hoomd
    from hoomd_blue_b200 import _lib
    from tests.mgpu_check import build
    fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "c1_64k.npz")))
    sim, cell, lj = build(hoomd, fx, hoomd.device.GPU(), thermostat=hoomd.md.methods.thermostats.Bussi(kT=1.5))
    sim.run(80)
    ke = sim.operations.integrator._thermo()[0]
    p2p = int(_lib.lib().hb2_domain_p2p_active(sim.state._h))
    kes = [None] * dist.get_world_size()
    dist.all_gather_object(kes, (ke, p2p, cell.num_builds))
    assert all(np.isfinite(k[0]) for k in kes), kes
    assert len({k[0] for k in kes}) == 1 and len({k[2] for k in kes}) == 1, kes
    T = 2.0 * ke / (3 * len(fx["pos"]) - 3)
    assert 1.3 < T < 1.7, T
    if dist.get_rank() == 0:
        print(f"p2p8 ok: world={dist.get_world_size()} grid={sim.state.domain_decomposition} p2p_mailboxes={p2p} "
              f"kT={T:.4f} builds={cell.num_builds}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
