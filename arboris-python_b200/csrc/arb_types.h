// Device-side view of a flattened model and of a batch's memory.
#pragma once
#include <stdint.h>
#include "../../include/arboris_b200.h"

#ifdef __CUDACC__
#define ARB_HDI __host__ __device__ __forceinline__
#else
#define ARB_HDI inline
#endif

ARB_HDI int arb_joint_ndof(int t) {
  return t == ARB_JOINT_FREE ? 6 : (t == ARB_JOINT_RZRYRX || t == ARB_JOINT_TXTYTZ) ? 3
         : (t == ARB_JOINT_RZRY || t == ARB_JOINT_RZRX || t == ARB_JOINT_RYRX) ? 2 : 1;
}
ARB_HDI int arb_joint_ngpos(int t) { return t == ARB_JOINT_FREE ? 16 : arb_joint_ndof(t); }
ARB_HDI int arb_cons_ndol(int t) {
  return t == ARB_CONS_JOINT_LIMITS ? 1 : t == ARB_CONS_BALL_SOCKET ? 3 : 4;
}

#define ARB_BODY_MASSIVE 1   /* some entry of the mass matrix > 0 (WeightController.init, controllers.py:37) */
#define ARB_BODY_HASMASS 2   /* mass matrix not identically zero */
#define ARB_BODY_HASVISC 4   /* viscosity matrix not identically zero */

// Read-only tables, device (or host, in the CPU unit-test build) pointers.
struct DevModel {
  int ndof, ngpos, nj, nc, na, nrows;
  int ncols;      // sum over bodies of the number of ancestor dofs (packed Jacobian columns)
  int maxk;       // longest root path, in dofs
  int anyvisc;
  const int *jtype, *jparent, *jdof, *jgpos;
  const double *Hpr, *HprInv, *Hcn, *HcnInv;   // [nj][12]  (R row-major, p)
  const int *hcn_ident;                        // [nj]
  const double *bmass, *bvisc;                 // [nj][36]
  const double *brx;                           // [nj][9]  M[0:3,3:6]/M[3,3] or 0 (core.py:1280-1283)
  const int *bflags;                           // [nj]
  const int *coloff;                           // [nj+1] first packed column of body b (b = 0 ground)
  const int *kcols;                            // [nj+1] number of path columns of body b
  const int *pathdof;                          // [ncols] dof of each packed column
  const int *ctype, *cint, *crow;              // constraints
  const double *cdbl;
  const int *atype, *aint;                     // controllers
  const double *adbl, *ablob;
  double up[3];
};

// Per-batch memory: caller-owned state + library-owned scratch, all [elem][W].
struct DevBatch {
  int64_t W;
  double *gpos, *gvel, *cforce;        // state (bound)
  // update_dynamic outputs
  double *pose;      // [nj][12]
  double *twist;     // [nj][6]
  double *J, *dJ;    // [ncols][6]
  double *M, *N, *B; // [n][n]
  // update_controllers outputs
  double *Z, *Y;     // [n][n]
  double *gforce;    // [n]
  // update_constraints
  double *cjac;      // [nrows][n]   active rows compacted
  double *cvel;      // [nrows]
  double *cA;        // [nrows][nrows]
  double *cT;        // [n][nrows]   Y J^T
  double *cpinv;     // [nrows][4]   pinv of each constraint's diagonal block
  double *caux;      // [nc][4]      sdist / pos0
  double *tmp;       // [2n]
  int *cactive, *cbranch, *cdol, *czidx /*[nc][3]*/;
  int *status;       // [W]
};
