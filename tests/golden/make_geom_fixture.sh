#!/bin/sh
# Regenerates the .geom fixture from the unmodified reference (needs /root/reference).
set -e
here=$(cd "$(dirname "$0")" && pwd)
g++ -std=c++17 -fopenmp -O1 -DNDEBUG -I/root/reference/include "$here/make_geom_fixture.cpp" -o /tmp/make_geom_fixture
/tmp/make_geom_fixture "$here/torus_patch.geom" "$here/torus_patch_coord_f64.bin" "$here/torus_patch_nodes_f64.bin"
