"""Host-side builders of the general element blocks (BASELINE.json configs 3-5) on inline cube meshes: DOF tables and
vertex coordinates as the reference's ConnManager / DOFManager / workset builder would hand them over.

Where the reference cannot produce the numbering itself (Q2 edge / face / cell ids come from STK-generated subcell
entities, Panzer_STKConnManager.cpp:160-226; mixed topologies throw in GeometricAggFieldPattern)
the Cartesian rule of dof-mgr/test/cartesian_topology/CartesianConnManager.cpp is used: DOFs are numbered by their
position on a refined lattice, lexicographically.  numpy only; the CUDA library does the assembly."""
from __future__ import annotations

import numpy as np

HEX8 = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]])
HEX27 = np.array([[-1,-1,-1],[1,-1,-1],[1,1,-1],[-1,1,-1],[-1,-1,1],[1,-1,1],[1,1,1],[-1,1,1],[0,-1,-1],[1,0,-1],[0,1,-1],[-1,0,-1],
                  [-1,-1,0],[1,-1,0],[1,1,0],[-1,1,0],[0,-1,1],[1,0,1],[0,1,1],[-1,0,1],[0,0,0],[0,0,-1],[0,0,1],[-1,0,0],[1,0,0],[0,-1,0],[0,1,0]])
HEX_EDGE = np.array([[0, 1], [1, 2], [2, 3], [3, 0], [4, 5], [5, 6], [6, 7], [7, 4], [0, 4], [1, 5], [2, 6], [3, 7]])
TET_EDGE = np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]])


def hex_cells(n, x_range=None):
    """(i, j, k) of the cells of an nx x ny x nz inline mesh, x fastest (element id order of CubeHexMeshFactory)"""
    nx, ny, nz = (n, n, n) if isinstance(n, int) else n
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ijk = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1)
    if x_range is not None:
        ijk = ijk[(ijk[:, 0] >= x_range[0]) & (ijk[:, 0] < x_range[1])]
    return ijk, (nx, ny, nz)


def hex_vertex_coords(ijk, dims, box=(1.0, 1.0, 1.0)):
    """cell_vertex_coordinates [ne][8][3] in Shards Hexahedron<8> order"""
    h = np.asarray(box, float) / np.asarray(dims, float)
    return (ijk[:, None, :] + (HEX8[None, :, :] + 1) // 2) * h[None, None, :]


def q1_node_lids(ijk, dims):
    nx, ny, _ = dims
    p = ijk[:, None, :] + (HEX8[None, :, :] + 1) // 2
    return (p[..., 0] + (nx + 1) * (p[..., 1] + (ny + 1) * p[..., 2])).astype(np.int32)


def three_field_lids(node_lids):
    """FieldAggPattern's interleaving of three nodal fields (dof-mgr/src/Panzer_FieldAggPattern.cpp:201-276)"""
    return (3 * node_lids[:, :, None] + np.arange(3)[None, None, :]).reshape(len(node_lids), -1).astype(np.int32)


def q2_hex_lids(ijk, dims):
    """27 DOFs per cell on the (2n+1)^3 lattice, Shards Hexahedron<27> node order"""
    nx, ny, _ = dims
    p = 2 * ijk[:, None, :] + 1 + HEX27[None, :, :]
    return (p[..., 0] + (2 * nx + 1) * (p[..., 1] + (2 * ny + 1) * p[..., 2])).astype(np.int64)


def hcurl_hex_lids(ijk, dims):
    """12 edge DOFs per cell (x-edges, y-edges, z-edges, each lexicographic) and the orientation signs
    (+1: the edge's first Shards vertex has the smaller global vertex id, Panzer_IntrepidOrientation.cpp:96-99)"""
    nx, ny, nz = dims
    a, b = HEX8[HEX_EDGE[:, 0]], HEX8[HEX_EDGE[:, 1]]
    direc = np.argmax(a != b, axis=1)
    lo = ijk[:, None, :] + ((a[None, :, :] > 0) & (np.arange(3)[None, None, :] != direc[None, :, None]))
    nxe = nx * (ny + 1) * (nz + 1); nye = (nx + 1) * ny * (nz + 1)
    idx = np.where(direc[None, :] == 0, lo[..., 0] + nx * (lo[..., 1] + (ny + 1) * lo[..., 2]),
          np.where(direc[None, :] == 1, nxe + lo[..., 0] + (nx + 1) * (lo[..., 1] + ny * lo[..., 2]),
                   nxe + nye + lo[..., 0] + (nx + 1) * (lo[..., 1] + (ny + 1) * lo[..., 2])))
    signs = np.where(b[np.arange(12), direc] > a[np.arange(12), direc], 1, -1).astype(np.int8)
    return idx.astype(np.int32), np.broadcast_to(signs[None, :], idx.shape).copy(), nxe + nye + (nx + 1) * (ny + 1) * nz


def cube_tets(ijk, dims, box=(1.0, 1.0, 1.0)):
    """CubeTetMeshFactory's split (Panzer_STK_CubeTetMeshFactory.cpp:397-465): 12 tetrahedra per hexahedron around its
    centroid.  Returns vertex positions on the (4n+1)^3 quarter lattice [ne][4][3] (integers) and coordinates."""
    faces = [(0, 1, 2, 3), (4, 7, 6, 5), (0, 4, 5, 1), (1, 5, 6, 2), (2, 6, 7, 3), (3, 7, 4, 0)]
    v = 4 * (ijk[:, None, :] + (HEX8[None, :, :] + 1) // 2)                  # hex vertices on the quarter lattice
    cen = 4 * ijk + 2
    tets = []
    for (p, q, r, s) in faces:
        for tri in ((p, q, r), (p, r, s)):
            tets.append(np.stack([v[:, tri[0]], v[:, tri[1]], v[:, tri[2]], cen], axis=1))
    t = np.stack(tets, axis=1).reshape(-1, 4, 3)
    h = np.asarray(box, float) / np.asarray(dims, float) / 4.0
    xyz = t * h[None, None, :]
    neg = np.linalg.det(xyz[:, 1:] - xyz[:, :1]) < 0
    t[neg] = t[neg][:, [0, 2, 1, 3]]; xyz[neg] = xyz[neg][:, [0, 2, 1, 3]]
    return t, xyz


def p2_tet_positions(tv):
    """10 node positions per tetrahedron on the quarter lattice doubled (vertex + mid-edge), Shards Tetrahedron<10> order"""
    mids = tv[:, TET_EDGE[:, 0]] + tv[:, TET_EDGE[:, 1]]
    return np.concatenate([2 * tv, mids], axis=1)


def compress(*position_arrays, dims):
    """lattice positions (ints, any common refinement) -> consecutive LIDs shared across the arrays; returns the LID
    arrays and the number of DOFs"""
    M = int(max(p.max() for p in position_arrays)) + 1
    keys = [p[..., 0].astype(np.int64) + M * (p[..., 1].astype(np.int64) + M * p[..., 2].astype(np.int64)) for p in position_arrays]
    uniq, inv = np.unique(np.concatenate([k.ravel() for k in keys]), return_inverse=True)
    out, at = [], 0
    for k in keys:
        out.append(inv[at:at + k.size].reshape(k.shape).astype(np.int32)); at += k.size
    return out, len(uniq)
