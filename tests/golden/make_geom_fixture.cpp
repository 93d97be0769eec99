// Generates tests/golden/torus_patch.geom with the UNMODIFIED reference (QuadElemList<double>::Write,
// include/sctl/experimental/quad_element.cpp:292-338) plus the arrays it was built from and what the reference reads
// back (QuadElemList::Read :340-397, GetNodeCoord), in SCTL's binary Vector format.  Run by make_geom_fixture.sh in the
// container that has /root/reference; only the outputs are committed.
#include <sctl.hpp>
#include <sctl/experimental/quad_element.hpp>
#include <sctl/experimental/quad_element.cpp>
using namespace sctl;
int main(int argc, char** argv) {
  const Integer order = 4;  // 3 order-4 elements on a torus patch, nodes at the reference's own parametric nodes
  const auto& nds = QuadElemList<double>::ParamNodes(order);
  const Long Nelem = 3, nn = order * order;
  Vector<double> coord(Nelem * nn * 3);  // AoS (element, node, xyz): the layout Init takes (:21-40)
  for (Long e = 0; e < Nelem; e++)
    for (Long i = 0; i < order; i++)
      for (Long j = 0; j < order; j++) {
        const double u = (e + nds[i]) * 0.4, v = nds[j] * 0.7;
        const double x[3] = {(1 + 0.25 * cos(v)) * cos(u), (1 + 0.25 * cos(v)) * sin(u), 0.25 * sin(v)};
        for (Integer k = 0; k < 3; k++) coord[(e * nn + i * order + j) * 3 + k] = x[k];
      }
  QuadElemList<double> lst;
  lst.Init(order, coord);
  lst.Write(argv[1]);
  coord.Write(argv[2]);
  QuadElemList<double> back;
  back.Read<double>(argv[1]);
  Vector<double> X;
  back.GetNodeCoord(&X, nullptr, nullptr);
  X.Write(argv[3]);
  return 0;
}
