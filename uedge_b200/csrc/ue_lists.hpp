// uedge_b200/csrc/ue_lists.hpp — host-side list building shared by the product library (ue_gpu.cu) and the CPU logic
// check of the kernels (tests/hostcheck).  Plain C++, no CUDA; include after ue_device.cuh (uses Win / make_win).
#pragma once
#include <algorithm>
#include <vector>

#include "ue_param_store.hpp"

// Candidate rows of a perturbation at cell (xc,yc): a superset of the cells whose residual rows can change.
// The four private cells are C0, Cw = ixm1(C0), Ce = ixp1(C0) (row yc connectivity) and Cs = (xc,yc-1).  A row
// (ix,iy') can read them only if iy' is within one row of yc and ix is within one poloidal step of {xw,xc,xe},
// where "one step" is taken through the index maps of the rows involved (and plain ix+-1), so that cells across an
// X-point cut are found where the maps connect them.  On a regular part of the mesh this is the 5 x 3 rectangle
// around (xc,yc).  With the integrated core-power condition (iflcore=1, boundary.m:485-523) the row that carries
// the poloidal sum, (min(ixpt2,nx), 0), also depends on every cell of rows 0 and 1.
inline void cell_candidates(const UeParams& P, int xc, int yc, std::vector<int>& out) {
  const int nxs = (int)P.nx + 2, nys = (int)P.ny + 2;
  auto M1 = [&](int ix, int iy) { return (int)P.ixm1[ix + nxs * iy]; };
  auto P1 = [&](int ix, int iy) { return (int)P.ixp1[ix + nxs * iy]; };
  const int seeds[3] = {M1(xc, yc), xc, P1(xc, yc)};
  out.clear();
  for (int iy = std::max(0, yc - 1); iy <= std::min(nys - 1, yc + 1); ++iy) {
    std::vector<char> in(nxs, 0);
    for (int sd : seeds) {
      in[sd] = 1;
      if (sd - 1 >= 0) in[sd - 1] = 1;
      if (sd + 1 < nxs) in[sd + 1] = 1;
      for (int r = std::max(0, std::min(iy, yc) - 1); r <= std::min(nys - 1, std::max(iy, yc) + 1); ++r) {
        in[M1(sd, r)] = 1; in[P1(sd, r)] = 1;
        for (int ix = 0; ix < nxs; ++ix) if (M1(ix, r) == sd || P1(ix, r) == sd) in[ix] = 1;
      }
    }
    for (int ix = 0; ix < nxs; ++ix) if (in[ix]) out.push_back(ix + nxs * iy);
  }
  // extrapolation boundary conditions (istepfc/istipfc/istewc/istiwc = 2) read the second interior row
  // (boundary.m:555-559, 1320-1324): a perturbation there changes the guard row two rows away
  // (likewise isnwconi/o = 2, and on the core boundary isupcore = 2 and isngcore = 3)
  {
    const bool core = P.isixcore[xc] == 1;
    const bool lo = core ? (P.isupcore == 2 || P.isngcore == 3) : (P.istepfcix[xc] == 2 || P.istipfcix[xc] == 2 || P.isnwconiix[xc] == 2);
    const bool hi = P.istewcix[xc] == 2 || P.istiwcix[xc] == 2 || P.isnwconoix[xc] == 2;
    if (yc == 2 && lo) out.push_back(xc);
    if (yc == nys - 3 && hi) out.push_back(xc + nxs * (nys - 1));
  }
  // half-space problem with a core region: a window that recomputes the electron-energy rows on both sides of the cut
  // face forms vex there from the zeroed upi (see f_upe_pre in ue_device.cuh), so those rows differ from yldot00 in
  // every such window: the cells (ixpt2, iy) and (ixpt2+1, iy), iy <= iysptrx1, inside the row window are candidates
  if (P.isfixlb == 2 && P.iysptrx1 > 0) {
    const Win w = make_win(P, xc, yc);  // ue_device.cuh (included before this header)
    const int ixc = (int)P.ixpt2;
    for (int iy = std::max(1, w.j2); iy <= std::min(w.j5, (int)P.iysptrx1); ++iy)
      for (int ix : {ixc, ixc + 1})
        if (ix >= w.i2 && ix <= w.i5 && std::find(out.begin(), out.end(), ix + nxs * iy) == out.end()) out.push_back(ix + nxs * iy);
  }
  if (P.iflcore == 1 && yc <= 1) {
    const int cell = std::min((int)P.ixpt2, (int)P.nx);  // row 0
    if (std::find(out.begin(), out.end(), cell) == out.end()) out.push_back(cell);
  }
  std::sort(out.begin(), out.end());
}

