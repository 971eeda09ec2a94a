"""AssemblyEngine-shaped driver over the C ABI (the call a user of the reference makes).

Mirrors disc-fe/src/Panzer_AssemblyEngine.hpp:68-127 (AssemblyEngine<EvalT>::evaluate, EvaluationFlags),
Panzer_AssemblyEngine_InArgs.hpp:92-107 (AssemblyEngineInArgs) and lof/Panzer_LinearObjContainer.hpp:63-72
(container members x, dxdt, d2xdt2, f, A).  The ghosted and the global container share storage: the
ghosted vectors/matrix are laid out owned ++ ghosted, so the global objects are their owned prefix
(DESIGN.md "containers").  torch is used only for device memory and for torch.distributed.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi, host


class EvaluationFlags:
    """AssemblyEngine<EvalT>::EvaluationFlags (Panzer_AssemblyEngine.hpp:72-86)"""
    Initialize, VolumetricFill, BoundaryFill, Scatter, All = 1, 2, 4, 8, 15

    def __init__(self, flags):
        if not (0 < flags <= EvaluationFlags.All):
            raise ValueError("EvaluationFlags: flags>0 && flags <= All")      # TEUCHOS_ASSERT at :74
        self.value = flags

    def getValue(self):
        return self.value


@dataclass
class LinearObjContainer:
    """TpetraLinearObjContainer: raw local views (device tensors or host arrays)."""
    x: object = None
    dxdt: object = None
    d2xdt2: object = None
    f: object = None
    A: object = None          # CSR values over the fill graph


@dataclass
class AssemblyEngineInArgs:
    """panzer::AssemblyEngineInArgs; alpha/beta start as NaN like the reference ("loud" defaults)."""
    ghostedContainer_: LinearObjContainer = None
    container_: LinearObjContainer = None
    alpha: float = float("nan")
    beta: float = float("nan")
    gamma: float = 0.0            # W_x_dot_dot_coeff (extension, see txasm.h)
    time: float = float("nan")
    step_size: float = float("nan")
    stage_number: float = 1.0
    gather_seeds: list = field(default_factory=list)
    evaluate_transient_terms: bool = False


class AssemblyEngine:
    """One engine per evaluation type, like AssemblyEngine_TemplateManager hands out."""

    def __init__(self, handle: capi.Handle, eval_type: int):
        self.h, self.eval_type = handle, eval_type

    def evaluate(self, inargs: AssemblyEngineInArgs, flags=EvaluationFlags.All):
        if isinstance(flags, EvaluationFlags):
            flags = flags.getValue()
        EvaluationFlags(flags)
        g = inargs.ghostedContainer_
        if g is None:
            raise ValueError("AssemblyEngineInArgs.ghostedContainer_ is null")
        if self.eval_type == capi.JACOBIAN and g.A is None:
            raise ValueError("Jacobian evaluation needs ghostedContainer_.A")
        self.h.evaluate(self.eval_type, g.x, g.f, g.A if self.eval_type == capi.JACOBIAN else None,
                        xdot=g.dxdt, xdotdot=g.d2xdt2, flags=flags, alpha=inargs.alpha, beta=inargs.beta,
                        gamma=inargs.gamma, time=inargs.time, zero_outputs=1,
                        evaluate_transient_terms=inargs.evaluate_transient_terms or g.dxdt is not None,
                        gather_seeds=inargs.gather_seeds)


# --------------------------------------------------------------------------------------------
@dataclass
class PoissonProblem:
    """Everything main.cpp of the Poisson example builds before its first evaluate
    (adapters-stk/example/PoissonExample/main.cpp:151-345), for the 3-D inline cube."""
    mesh: host.Mesh
    dof: host.DOFManager
    handle: capi.Handle
    n_cells: int
    n_owned: int
    n_local: int
    nnz: int
    plan: dict = None
    dirichlet_dofs: np.ndarray = None
    keep: list = field(default_factory=list)


def build_poisson_problem(n, rank=0, nranks=1, comm=None, procs=(-1, -1, -1), perturb=0.0, device=0,
                          scatter_mode=capi.SCATTER_AUTO, terms=None, dirichlet=True, nccl_uid=None, stream=None, p2p=True,
                          device_setup=None):
    """Mesh -> connectivity -> DOFManager -> graph -> txasm handle (+ Dirichlet nodesets + halo plan).

    n: int or (nx, ny, nz) GLOBAL element counts.  comm: host.TorchComm for nranks > 1.
    device_setup: build the mesh tables and the DOF numbering on the device (device_setup.py) instead of
    through the host mirror.  Default: yes for unperturbed meshes (the perturbation rule lives in the host mirror) on 1 or 2
    ranks -- the configurations run on B200s over NCCL so far; more ranks are covered by the gloo tests only, so there it is
    opt-in (TXASM_DEVICE_SETUP=1) until it has run on hardware.  TXASM_HOST_SETUP=1 forces the host mirror.  For nranks > 1
    the default torch.distributed group moves the directory exchanges.
    """
    import os
    import torch
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    fac = host.CubeHexMeshFactory(**{"X Elements": nx, "Y Elements": ny, "Z Elements": nz,
                                     "X Procs": procs[0], "Y Procs": procs[1], "Z Procs": procs[2]})
    dev = torch.device(f"cuda:{device}")
    if device_setup is None:
        device_setup = (not perturb and os.environ.get("TXASM_HOST_SETUP") != "1" and
                        (nranks <= 2 or os.environ.get("TXASM_DEVICE_SETUP") == "1"))
    if device_setup:
        from .device_setup import DeviceDOFManager, DeviceMesh
        mesh = DeviceMesh(fac, rank, nranks, device=dev)
        if perturb:
            mesh.perturb(perturb)
        dof = DeviceDOFManager(rank, nranks)
        dof.setConnManager(mesh.connectivity_t())
        dof.addField("TEMPERATURE")
        dof.buildGlobalUnknowns()
        lids = dof.lids_t()
        cc = mesh.cell_vertex_coordinates_t()
        dof_host = dof.host_manager() if nranks > 1 else None
    else:
        mesh = fac.buildMesh(rank, nranks)
        if perturb:
            mesh.perturb(perturb)
        dof = host.DOFManager(rank, nranks)
        dof.setConnManager(mesh.getConnectivity())
        dof.addField("TEMPERATURE")
        dof.buildGlobalUnknowns(comm)
        lids = torch.from_numpy(dof.getLIDs()).to(dev)
        cc = torch.from_numpy(mesh.cell_vertex_coordinates()).to(dev)
        dof_host = dof
    h = capi.Handle(device=device, stream=stream, scatter_mode=scatter_mode)
    h.block_add(lids, cell_coords=cc, n_rows=dof.num_local)
    del cc
    if device_setup:
        mesh.release_coordinates()
    plan = None
    keep = [lids]
    if nranks == 1:
        nnz = h.graph_build()                 # buildGhostedGraph on the device
    else:
        h.graph_build()                       # buildGhostedGraph on the device
        # Import/Export negotiation on the host from the ghost rows alone (surface size); the fill graph -- owned rows gain
        # the columns other ranks contribute (buildGraph's Export INSERT) -- is merged on the device, which also returns
        # the position of every matrix value this rank will receive.  The 2 GB graph never visits the host.
        no, nl = dof.num_owned, dof.num_local
        g_rp, g_ci = h.graph_get_rows(no, nl - no)
        lof = host.TpetraLinearObjFactory(dof_host)
        lof.setGhostRows(g_rp, g_ci)
        lof.buildPlans(comm)
        plan = lof.plan()
        plan["mat_recv_pos"], nnz = h.graph_merge_columns(plan["pair_rows"], plan["pair_cols"])
        if nccl_uid is not None:
            h.comm_init(nranks, rank, nccl_uid)
            h.halo_set(dof.num_owned, plan["nbr_rank"], plan["send_off"], plan["send_lids"], plan["recv_off"], plan["recv_lids"])
            h.halo_set_matrix(plan["mat_recv_off"], plan["mat_recv_pos"])
            if p2p:
                # peer-memory exchange: publish my receive slab, map the neighbours' (cudaIpc handles travel through
                # the host communicator)
                import torch.distributed as dist
                blobs = [None] * nranks
                dist.all_gather_object(blobs, h.halo_p2p_export())
                h.halo_p2p_connect(blobs)
                dist.barrier()
    h.terms_set(terms if terms is not None else capi.poisson_terms())
    ddofs = None
    if dirichlet:
        # the six side sets, value 0 (3-D analogue of PoissonExample/main.cpp:153-180); a node shared by
        # several sides appears once.  Node id -> LID through the element tables.
        # Only cells on the boundary of the global brick carry such nodes: select them by element id first
        # (id - 1 = ix + NX (iy + NY iz), Panzer_STK_CubeHexMeshFactory.cpp:446), then their vertices by node id
        # (id - 1 = I + (NX+1) (J + (NY+1) K), :401-461).
        if device_setup:                              # same selection with device arrays
            eid = mesh.elem_ids_t() - 1
            ix, iy, iz = eid % nx, (eid // nx) % ny, eid // (nx * ny)
            bc = torch.nonzero((ix == 0) | (ix == nx - 1) | (iy == 0) | (iy == ny - 1) | (iz == 0) | (iz == nz - 1)).reshape(-1)
            nid = mesh.elem_nodes_t()[bc] - 1
            I, J, K = nid % (nx + 1), (nid // (nx + 1)) % (ny + 1), nid // ((nx + 1) * (ny + 1))
            on = (I == 0) | (I == nx) | (J == 0) | (J == ny) | (K == 0) | (K == nz)
            ddofs = torch.unique(lids[bc][on]).cpu().numpy().astype(np.int32)
        else:
            eid = mesh.elem_ids() - 1
            ix, iy, iz = eid % nx, (eid // nx) % ny, eid // (nx * ny)
            bc = np.nonzero((ix == 0) | (ix == nx - 1) | (iy == 0) | (iy == ny - 1) | (iz == 0) | (iz == nz - 1))[0]
            nid = mesh.elem_nodes()[bc] - 1
            I, J, K = nid % (nx + 1), (nid // (nx + 1)) % (ny + 1), nid // ((nx + 1) * (ny + 1))
            on = (I == 0) | (I == nx) | (J == 0) | (J == ny) | (K == 0) | (K == nz)
            ddofs = np.unique(dof.getLIDs()[bc][on]).astype(np.int32)
        h.dirichlet_set(ddofs, np.zeros(len(ddofs)))
    h.setup()
    return PoissonProblem(mesh, dof, h, mesh.num_elems, dof.num_owned, dof.num_local, nnz, plan, ddofs, keep)
