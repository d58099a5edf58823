"""Regenerate the committed mesh fixtures from the reference's gmsh files.

Run in the build container only (``/root/reference`` does not exist on the GPU box):
    python tests/golden/make_fixtures.py
The reference cannot be imported here (no FEniCS), so the only reference-made artefacts
available are its gmsh 4.1 ASCII mesh fixtures; they are converted with the oracle's own
reader into small ``.npz`` files (coords, cells, boundary facets + physical tags).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import meshes  # noqa: E402

REF = "/root/reference/tests/mesh"
for name, rel in (
    ("unit_circle", "unit_circle/mesh.msh"),  # == demos/documented/shape_optimization/shape_poisson/mesh/mesh.msh
    ("mesh2d", "mesh.msh"),
    ("mesh3d", "mesh3.msh"),
):
    m = meshes.read_gmsh41(os.path.join(REF, rel))
    meshes.save_npz(m, os.path.join(HERE, name + ".npz"))
    print(name, m.num_vertices, m.num_cells, m.facets.shape[0], sorted(set(m.facet_tags.tolist())))
