"""``MeshGenerator`` with the reference's call surface (``datagen/mesh_generator.py:58-521``:
``generate_geometry``, ``normalize_geometry``, ``generate_mesh``, ``sample_conditions``) on top of
the in-house seeded plate generator.  The reference uses shapely + gmsh + sklearn, which this
image does not have; meshing stays on the host, outside the timed path, either way."""
from __future__ import annotations

import numpy as np

from ..plates import GeometryRejected, PlateGenerator


def write_medit(filepath: str, coors: np.ndarray, conn: np.ndarray) -> None:
    """MEDIT ``.mesh`` text as gmsh writes it (``Dimension 3``, z = 0, 1-based cells, reference 0)
    with 17 significant digits so that every double survives the round trip bit for bit."""
    k = conn.shape[1]
    name = {3: "Triangles", 4: "Quadrilaterals"}[k]
    with open(filepath, "w") as f:
        f.write(" MeshVersionFormatted 2\n Dimension\n 3\n Vertices\n %d\n" % len(coors))
        f.write("".join("%.17g %.17g 0 0\n" % (x, y) for x, y in coors))
        f.write(" %s\n %d\n" % (name, len(conn)))
        f.write("".join(" ".join(str(v + 1) for v in row) + " 0\n" for row in conn))
        f.write(" End\n")


class MeshGenerator(PlateGenerator):
    def generate_geometry(self):
        for _ in range(200):
            try:
                return super().generate_geometry()
            except GeometryRejected:
                continue
        raise GeometryRejected("no valid geometry in 200 draws")

    def generate_mesh(self, geometry, filename="part", mesh_size: float = 1e-2, view_mesh: bool = False):
        """Meshes ``geometry``, writes ``<filename>.mesh`` and returns (polygons_ptags,
        polygons_ltag_ptags) like the reference (``mesh_generator.py:246-317``)."""
        ptags, ltags = super().generate_mesh(geometry, mesh_size)
        coors, conn = self.mesh
        write_medit(filename + ".mesh", coors, conn)
        return ptags, ltags
