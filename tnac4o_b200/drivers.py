"""Thin drivers mirroring the reference's example scripts (examples/e01 ... e06): one function per script, same argument
meaning, built on the solver object of this package.  They take a coupling list (or an RMF model) instead of a file name
where the reference reads ``./../instances/...`` -- file loading is :func:`tnac4o_b200.load_Jij`.

The only driver with logic of its own is :func:`search_gs_degeneracy` (examples/e06_search_gs_degeneracy_J124.py:98-110):
four searches from the four edges of the lattice (independent -> run as concurrent replicas on one GPU, or spread over the
ranks of a process group), then the reference's selection rule: lowest energy, and among the rotations that reach it the
largest degeneracy.
"""
import os

import numpy as np

from .solver import tnac4o, load


def search_gs(J, Nx, Ny, Nc=8, rot=0, beta=3, D=48, M=2 ** 10, relative_P_cutoff=1e-8, precondition=True, device=None):
    """examples/e01_search_gs_droplet_instances.py:22-80 (and e06:23-75 with its own defaults)"""
    ins = tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=Nc, J=J, beta=beta, device=device)
    if rot > 0:
        ins.rotate_graph(rot=rot)
    if precondition:
        ins.precondition(mode='balancing')
    ins.search_ground_state(M=M, relative_P_cutoff=relative_P_cutoff, Dmax=D)
    return ins


def gibbs_sampling(J, Nx, Ny, Nc=8, rot=0, beta=1, D=48, M=2 ** 10, precondition=True, device=None):
    """examples/e02_sample_droplet_instances.py"""
    ins = tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=Nc, J=J, beta=beta, device=device)
    if rot > 0:
        ins.rotate_graph(rot=rot)
    if precondition:
        ins.precondition(mode='balancing')
    ins.gibbs_sampling(M=M, Dmax=D)
    return ins


def search_spectrum(J, Nx, Ny, Nc=8, rot=0, beta=3, D=48, M=2 ** 10, relative_P_cutoff=1e-8, excitations_encoding=1, dE=1.0,
                    hd=0, precondition=True, device=None):
    """examples/e03_search_spectrum_droplet_instances.py (noise of 1e-7 for the adjacency encodings, as there)"""
    ins = tnac4o(mode='Ising', Nx=Nx, Ny=Ny, Nc=Nc, J=J, beta=beta, device=device)
    if rot > 0:
        ins.rotate_graph(rot=rot)
    if excitations_encoding > 1:
        ins.add_noise(amplitude=1e-7)
    if precondition:
        ins.precondition(mode='balancing')
    ins.search_low_energy_spectrum(excitations_encoding=excitations_encoding, M=M, relative_P_cutoff=relative_P_cutoff, Dmax=D,
                                   max_dEng=dE, lim_hd=hd)
    return ins


def load_spectrum(file_name, J=None, dE=1.0, max_states=2 ** 20):
    """examples/e04_load_spectrum_droplet_instances.py: load, decode, and (when the couplings are given) the consistency
    error between the decoded energies and energy_Jij of the decoded states"""
    from .auxx import energy_Jij
    ins = load(file_name)
    ins.decode_low_energy_states(max_dEng=dE, max_states=max_states)
    error = None
    if J is not None:
        error = float(np.max(np.abs(ins.energy - energy_Jij(J, ins.binary_states()))))
    return ins, error


def write_states_txt(ins, file_name):
    """the text format of examples/e02:119-131: one line per state, energy first, then the spins (1 up, 0 down)"""
    bits = ins.binary_states()
    with open(file_name, 'w') as f:
        print("# One line per state; First column is the energy, the rest is a state; \
                1 = spin up = si=+1; 0 = spin down = si=-1", file=f)
        for e, row in zip(ins.energy, bits):
            line = np.zeros((1, ins.L + 1))
            line[0, 0], line[0, 1:] = e, row
            np.savetxt(f, line, fmt=' '.join(['%4.6f'] + ['%i'] * ins.L), delimiter=' ')


def search_gs_degeneracy(J, Nx, Ny, Nc=8, beta=0.75, D=48, M=2 ** 12, relative_P_cutoff=1e-8, precondition=True,
                         rotations=(0, 1, 2, 3), concurrent=True, device=None, group=None):
    """examples/e06:98-110.  Returns (energy, degeneracy, per_rotation) with per_rotation = [(rot, energy, degeneracy), ...].
    The four rotations are independent searches: on one GPU they run as concurrent replicas (one host thread and CUDA
    stream each); with a torch.distributed group every rank takes the rotations ``rank::world`` and the (energy,
    degeneracy) pairs are exchanged at the end -- no data-path collective."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if (dist.is_available() and dist.is_initialized()) else (0, 1)
    mine = [r for i, r in enumerate(rotations) if i % world == rank]

    def one(rot):
        ins = search_gs(J, Nx, Ny, Nc=Nc, rot=rot, beta=beta, D=D, M=M, relative_P_cutoff=relative_P_cutoff,
                        precondition=precondition, device=device)
        return (int(rot), float(ins.energy[0]), int(ins.degeneracy))

    if concurrent and len(mine) > 1:
        from .parallel import run_concurrently
        dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        found = run_concurrently([(lambda r=r: one(r)) for r in mine], device=dev)
    else:
        found = [one(r) for r in mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, found, group=group)
        found = sorted(x for p in parts for x in p)
    energy = min(e for _, e, _ in found)
    degeneracy = max(d for _, e, d in found if e == energy)
    return energy, degeneracy, found


def write_gs_degeneracy_txt(file_name, energy, degeneracy):
    """examples/e06:113-120"""
    np.savetxt(file_name, np.array([energy, degeneracy], dtype=int), delimiter=' ', header='Energy and degeneracy', fmt='%1d')


def minimal_RMF(J, Nx, Ny, rot=0, beta=4, D=32, M=1024, relative_P_cutoff=1e-12, excitations_encoding=1, dE=3.1, hd=0,
                max_states=100, precondition=False, device=None):
    """examples/e05_minimal_RMF.py:22-77 for a given RMF model dictionary"""
    ins = tnac4o(mode='RMF', Nx=Nx, Ny=Ny, J=J, beta=beta, device=device)
    if rot > 0:
        ins.rotate_graph(rot=rot)
    if excitations_encoding > 1:
        ins.add_noise(amplitude=1e-7)
    if precondition:
        ins.precondition(mode='balancing')
    ins.search_low_energy_spectrum(excitations_encoding=excitations_encoding, M=M, relative_P_cutoff=relative_P_cutoff, Dmax=D,
                                   max_dEng=dE, lim_hd=hd)
    ins.decode_low_energy_states(max_dEng=dE, max_states=max_states)
    return ins


def results_dir():
    d = os.path.join(os.getcwd(), 'results')
    os.makedirs(d, exist_ok=True)
    return d
