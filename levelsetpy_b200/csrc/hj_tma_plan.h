// hj_tma_plan.h -- host-side plan of the TMA backend: one tensor map per RK buffer + tile geometry.
#pragma once
#include <cuda.h>

#include "hj_tma_kernel.cuh"
#include "hj_vec_kernel.cuh"

struct HjTmaPlan {
  CUtensorMap tmap[3];     // haloed (TY+6) x (TX+8) boxes on the three RK buffers
  CUtensorMap tmap_y0;     // un-haloed TY x TX box on buffer 0 (y at the start of the step)
  TmaGeom geo;
  int tx, ty;
  long long nblocks;
  long long tiles;         // (X, Y, slow) tiles: nblocks = tiles * geo.nzc
  // dimension-split path (product systems): pass 2 tensor maps + geometry
  bool split = false;
  bool thin = false;      // pass 2 runs the thin-slab tile (SplitCfg::P2T)
  CUtensorMap vmap[3];
  VecGeom vgeo;
  long long vblocks = 0;
  int vb = 0;              // doubles per vector chunk of pass 2 (the column quantum of hj_stage_pass_cols)
  long long vlen = 0;      // length of the vector axis (flattened trailing dims, pitched)
};
