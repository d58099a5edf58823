"""Input / output of the device layer (``cashocs/io/__init__.py``): config files and gmsh meshes."""

from cashocs_b200.io.config import Config
from cashocs_b200.io.config import load_config
from cashocs_b200.io.mesh import read_gmsh
from cashocs_b200.io.mesh import write_out_mesh

__all__ = ["Config", "load_config", "read_gmsh", "write_out_mesh"]
