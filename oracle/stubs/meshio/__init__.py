"""Minimal stand-in for meshio, used ONLY to import the reference FEMpy package inside the
build container. Provides `read` for gmsh-4.1 ASCII files holding a single block of 4-node
quads (Examples/Meshes/LBracket.msh) and a do-nothing `Mesh`. Test infrastructure, not product.
"""
import numpy as np


class Mesh:
    def __init__(self, points=None, cells=None, point_data=None, cell_data=None, **kwargs):
        self.points = points
        self.cells = cells
        self.point_data = point_data
        self.cell_data = cell_data
        self.cells_dict = {}

    def write(self, *args, **kwargs):
        pass


_GMSH_TYPES = {1: ("line", 2), 2: ("triangle", 3), 3: ("quad", 4), 9: ("triangle6", 6), 10: ("quad9", 9), 15: ("vertex", 1)}


def read(path):
    """Read a gmsh 4.1 ASCII mesh: returns an object with .points (N,3) and .cells_dict."""
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    i = lines.index("$Nodes") + 1
    numBlocks, numNodes, _, _ = (int(v) for v in lines[i].split())
    i += 1
    points = np.zeros((numNodes, 3))
    tagToIndex = {}
    count = 0
    for _ in range(numBlocks):
        _, _, _, n = (int(v) for v in lines[i].split())
        i += 1
        tags = [int(lines[i + k]) for k in range(n)]
        i += n
        for k in range(n):
            points[count] = [float(v) for v in lines[i + k].split()]
            tagToIndex[tags[k]] = count
            count += 1
        i += n
    i = lines.index("$Elements") + 1
    numBlocks, numElems, _, _ = (int(v) for v in lines[i].split())
    i += 1
    cells = {}
    for _ in range(numBlocks):
        _, _, elType, n = (int(v) for v in lines[i].split())
        i += 1
        name, nn = _GMSH_TYPES[elType]
        block = np.array([[tagToIndex[int(v)] for v in lines[i + k].split()[1 : 1 + nn]] for k in range(n)], dtype=np.int64)
        i += n
        cells[name] = block if name not in cells else np.vstack((cells[name], block))
    mesh = Mesh(points=points)
    mesh.cells_dict = cells
    return mesh
