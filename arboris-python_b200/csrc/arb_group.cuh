// The prepare stage with a GROUP of 16 lanes per world (half a warp) and on-chip scratch.
//
// Same arithmetic as world_fused_prepare (arb_fused.cuh / arb_artic.cuh; reference
// core.py:682-734, 811-818, 910-927), organised so that the intermediates of the articulated
// factorisation never leave the SM:
//
//   * every world owns a record of ~24 KB of SHARED memory (GroupLayout, arb_types.h): X, s, s^,
//     U, LA, LM, 1/d, u of every joint / dof -- what the lane-per-world stage streams through HBM
//     (70 KB of DRAM traffic per world-step) -- plus small exchange buffers;
//   * kinematics: one lane per joint (closed forms of joints.py), then a root-to-leaf scan by
//     tree depth for poses, twists and theta;
//   * factorisation: lanes 0-5 own row r AND column r of the articulated matrix IA, lanes 8-13 of
//     IM.  U = IA s + IM s^ comes from the rows, s^T IA and s^T IM from the columns, the rank-one
//     updates and the congruence X^T . X are applied to both copies: no transposition, one
//     barrier per dof and two per joint;
//   * solves: one lane per right-hand side -- the free motion and the NG generator columns (unit
//     wrenches on the bodies that carry constraint frames, unit forces on limited dofs) run side by
//     side, each lane carrying its (V, V^) in registers; the rows they read are shared-memory
//     broadcasts;
//   * hand-over to the Gauss-Seidel and finish stages through HBM: q_free, v0, Lambda, the contact
//     maps, and K = Z^-1 G^T over ALL dofs, so that the finish stage is q'+ = q_free + K y instead
//     of another articulated solve over rows that no longer exist in HBM.
//
// The code below also compiles for the host (tests/hosttest), where the lanes of a group are run
// one after the other between the barriers: GRP_LANES ... GRP_SYNC brackets a region in which a
// lane only reads what earlier regions (or itself) wrote.
#pragma once
#include "arb_fused.cuh"

#define ARB_GL 16

struct GroupCtx {
  double* sm;       // this world's shared-memory record (GroupLayout)
  int lane;         // lane in the group (device)
  unsigned mask;    // the group's lanes in the warp (device)
};

#ifdef __CUDA_ARCH__
#define GRP_LANES(g) { const int lane = (g).lane;
#define GRP_SYNC(g) } __syncwarp((g).mask);
#define GRP_END(g) }
#define GRP_V(x) x
#define GRP_DECL(type, name) type name
#define GRP_DECLA(type, name, n) type name[n]
#else
#define GRP_LANES(g) for (int lane = 0; lane < ARB_GL; ++lane) {
#define GRP_SYNC(g) }
#define GRP_END(g) }
#define GRP_V(x) x[lane]
#define GRP_DECL(type, name) type name[ARB_GL]
#define GRP_DECLA(type, name, n) type name[ARB_GL][n]
#endif

ARB_D void grp_load_se3(const double* p, Se3& h) {
#pragma unroll
  for (int i = 0; i < 9; ++i) h.R[i] = p[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) h.p[i] = p[9 + i];
}
ARB_D void grp_store_se3(double* p, const Se3& h) {
#pragma unroll
  for (int i = 0; i < 9; ++i) p[i] = h.R[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) p[9 + i] = h.p[i];
}

// body term of the articulated matrices (as artic_factor): A_b = M_b/dt + B_b + Omega(T) M_b +
// M_b adjacency(theta) (36, row-major) and the bias -M_b (T/dt + gravity_b) (6) -> out[42]
ARB_D void grp_body_term(const DevModel& m, int j, const double* T, const double* th, const Se3& pose,
                         double idt, double* out) {
  const int flags = m.bflags[j];
  const double* Mb = m.bmass + 36 * j;
#pragma unroll
  for (int i = 0; i < 42; ++i) out[i] = 0.;
  if (flags & ARB_BODY_HASMASS) {
    double X3[9];
    {
      double wx[9], t1[9], t2[9];
      skew3(T, wx);
      m3_mul(m.brx + 9 * j, wx, t1);
      m3_mul(wx, m.brx + 9 * j, t2);
#pragma unroll
      for (int i = 0; i < 9; ++i) X3[i] = t1[i] - t2[i];
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {  // Omega M, column by column
      double top[3] = {Mb[c], Mb[6 + c], Mb[12 + c]}, bot[3] = {Mb[18 + c], Mb[24 + c], Mb[30 + c]};
      double a[3], x2[3], d[3];
      cross3(T, top, a);
      m3_mulv(X3, bot, x2);
      cross3(T, bot, d);
#pragma unroll
      for (int i = 0; i < 3; ++i) { out[6 * i + c] += a[i] + x2[i]; out[6 * (i + 3) + c] += d[i]; }
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {  // M adjacency(theta), row by row
      const double* a1 = Mb + 6 * r;
      const double* a2 = Mb + 6 * r + 3;
      double c1[3], c2[3], c3[3];
      cross3(a1, th, c1);
      cross3(a2, th + 3, c2);
      cross3(a2, th, c3);
#pragma unroll
      for (int i = 0; i < 3; ++i) { out[6 * r + i] += c1[i] + c2[i]; out[6 * r + 3 + i] += c3[i]; }
    }
#pragma unroll
    for (int i = 0; i < 36; ++i) out[i] += Mb[i] * idt;
    double a[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) a[i] = T[i] * idt;
    if ((flags & ARB_BODY_MASSIVE) && m.nweight > 0) {
      const double gt[6] = {0., 0., 0., m.gravity * m.up[0], m.gravity * m.up[1], m.gravity * m.up[2]};
      double g[6];
      iad_apply(pose, gt, g);
#pragma unroll
      for (int i = 0; i < 6; ++i) a[i] += g[i];
    }
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      double t = 0.;
#pragma unroll
      for (int c = 0; c < 6; ++c) t += Mb[6 * r + c] * a[c];
      out[36 + r] = -t;
    }
  }
  if (flags & ARB_BODY_HASVISC) {
    const double* Bb = m.bvisc + 36 * j;
#pragma unroll
    for (int i = 0; i < 36; ++i) out[i] += Bb[i];
  }
}

// `write_poses`: also leave body poses and twists in the HBM scratch (read-backs between the two
// halves of a step, arb_step_begin / arb_step_end).
ARB_D void group_prepare(const DevModel& m, const DevBatch& b, int64_t w, double dt, GroupCtx& g, bool write_poses) {
  const GroupLayout& L = m.gl;
  const int n = m.ndof, nj = m.nj, NG = m.ngrows;
  double* sX = g.sm + L.X;
  double* sS = g.sm + L.S;
  double* sSh = g.sm + L.Sh;
  double* sU = g.sm + L.U;
  double* sLA = g.sm + L.LA;
  double* sLM = g.sm + L.LM;
  double* sdinv = g.sm + L.dinv;
  double* su0 = g.sm + L.u0;
  double* sPose = g.sm + L.kin;              // [nj][12]
  double* sT = sPose + 12 * nj;              // [nj][6]
  double* sTh = sT + 6 * nj;                 // [nj][6]
  double* sSlot = g.sm + L.slot;             // [nslot][78]  (aliases the kinematics area)
  double* sUx = g.sm + L.ex;                 // [2][2][6]
  double* sLx = sUx + 24;                    // [2][2][6]
  double* sBx = sLx + 24;                    // [2][6]
  double* sBm = sBx + 12;                    // [2][36]
  double* sCm = sBm + 72;                    // [2][36]
  double* sFlag = sCm + 72;                  // [4]
  double* sAb = g.sm + L.ab;                 // [nj][42]
  double* sAu = g.sm + L.au;                 // [16][maxpath]
  double* sAv = g.sm + L.av;                 // [nvslot][16][12]
  double* sRe = g.sm + L.re;                 // [ngen][9]
  const double idt = 1. / dt;

  // ---- joint-local kinematics: one lane per joint (artic_kinematics, first half) --------------
  GRP_LANES(g)
    if (lane == 0) { sFlag[0] = 0.; sFlag[1] = 0.; }
    for (int j = lane; j < nj; j += ARB_GL) {
      const int type = m.jtype[j];
      const int nd = arb_joint_ndof(type);
      const int dof = m.jdof[j];
      double q[16], dq[6];
      for (int i = 0; i < arb_joint_ngpos(type); ++i) q[i] = ST_LD(b.gpos, m.jgpos[j] + i);
      for (int i = 0; i < nd; ++i) dq[i] = ST_LD(b.gvel, dof + i);
      JointKin k;
      joint_kinematics(type, q, dq, k);
      const bool ident = m.hcn_ident[j] != 0;
      Se3 Hpr, Hcn, Hpc, t0;
      load_se3_const(m.Hpr, j, Hpr);
      se3_mul(Hpr, k.H, t0);
      if (ident) {
        Hpc = t0;
        se3_identity(Hcn);
      } else {
        Se3 HcnInv;
        load_se3_const(m.HcnInv, j, HcnInv);
        load_se3_const(m.Hcn, j, Hcn);
        se3_mul(t0, HcnInv, Hpc);
      }
      grp_store_se3(sX + 12 * j, Hpc);
      // local parts of the twist and of theta: Ad_cn T_nr and Ad_cn Ad_nr Ad_nr T_nr
      double tb[6], tau[6];
      {
        Se3 Hnr;
        se3_inv(k.H, Hnr);
        double t1[6], t2[6];
        ad_apply(Hnr, k.T, t1);
        ad_apply(Hnr, t1, t2);
        if (ident) {
#pragma unroll
          for (int i = 0; i < 6; ++i) { tb[i] = k.T[i]; tau[i] = t2[i]; }
        } else {
          ad_apply(Hcn, k.T, tb);
          ad_apply(Hcn, t2, tau);
        }
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) { sT[6 * j + i] = tb[i]; sTh[6 * j + i] = tau[i]; }
      // own columns: s = Ad_cn S; Sh holds Ad_cn dS until theta is known
      for (int c = 0; c < nd; ++c) {
        double s0[6], ds0[6], s[6], ds[6];
        if (type == ARB_JOINT_FREE) {
#pragma unroll
          for (int i = 0; i < 6; ++i) { s0[i] = (i == c) ? 1. : 0.; ds0[i] = 0.; }
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) { s0[i] = k.S[6 * c + i]; ds0[i] = k.dS[6 * c + i]; }
        }
        if (ident) {
#pragma unroll
          for (int i = 0; i < 6; ++i) { s[i] = s0[i]; ds[i] = ds0[i]; }
        } else {
          ad_apply(Hcn, s0, s);
          ad_apply(Hcn, ds0, ds);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { sS[6 * (dof + c) + i] = s[i]; sSh[6 * (dof + c) + i] = ds[i]; }
      }
    }
  GRP_SYNC(g)

  // ---- root-to-leaf scan by depth: poses, twists, theta (core.py:1295-1308) --------------------
  for (int lv = 0; lv < L.nlev; ++lv) {
    GRP_LANES(g)
      for (int i = m.glev_off[lv] + lane; i < m.glev_off[lv + 1]; i += ARB_GL) {
        const int j = m.glev_joint[i];
        const int par = m.jparent[j];
        Se3 Hpc, Hgc;
        grp_load_se3(sX + 12 * j, Hpc);
        double T[6], th[6];
        if (par == 0) {
          Hgc = Hpc;
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) { T[k2] = sT[6 * j + k2]; th[k2] = 0. - sTh[6 * j + k2]; }
        } else {
          Se3 Hgp;
          grp_load_se3(sPose + 12 * (par - 1), Hgp);
          se3_mul(Hgp, Hpc, Hgc);
          double Tp[6], thp[6], ta[6], tc[6];
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) { Tp[k2] = sT[6 * (par - 1) + k2]; thp[k2] = sTh[6 * (par - 1) + k2]; }
          iad_apply(Hpc, Tp, ta);
          iad_apply(Hpc, thp, tc);
#pragma unroll
          for (int k2 = 0; k2 < 6; ++k2) { T[k2] = ta[k2] + sT[6 * j + k2]; th[k2] = tc[k2] - sTh[6 * j + k2]; }
        }
        grp_store_se3(sPose + 12 * j, Hgc);
#pragma unroll
        for (int k2 = 0; k2 < 6; ++k2) { sT[6 * j + k2] = T[k2]; sTh[6 * j + k2] = th[k2]; }
      }
    GRP_SYNC(g)
  }

  // ---- s^ = Ad_cn dS - adjacency(theta) s; body terms; constraint activation and maps -----------
  GRP_LANES(g)
    for (int k = lane; k < n; k += ARB_GL) {
      const int j = m.dofjoint[k];
      double th[6], s[6], as[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { th[i] = sTh[6 * j + i]; s[i] = sS[6 * k + i]; }
      adj_apply(th, s, as);
#pragma unroll
      for (int i = 0; i < 6; ++i) sSh[6 * k + i] -= as[i];
    }
    for (int j = lane; j < nj; j += ARB_GL) {
      double T[6], th[6], out[42];
      Se3 pose;
#pragma unroll
      for (int i = 0; i < 6; ++i) { T[i] = sT[6 * j + i]; th[i] = sTh[6 * j + i]; }
      grp_load_se3(sPose + 12 * j, pose);
      grp_body_term(m, j, T, th, pose, idt, out);
#pragma unroll
      for (int i = 0; i < 42; ++i) sAb[42 * j + i] = out[i];
      if (write_poses) {
#pragma unroll
        for (int i = 0; i < 12; ++i) FT(b.fpose, 12 * j + i) = sPose[12 * j + i];
#pragma unroll
        for (int i = 0; i < 6; ++i) FT(b.atw, 6 * j + i) = T[i];
      }
    }
    // frames of the contact-aligned generator bodies: R_e = R_c^T R_body
    for (int gi = lane; gi < m.ngen; gi += ARB_GL) {
      double Re[9];
      if (m.gen_aligned[gi]) {
        double Rc[9];
        int zi[3];
        zaligned(m.cdbl + ARB_CONS_NDBL * m.gen_c0[gi] + 32, Rc, zi);
        const double* Rb = sPose + 12 * (m.gen_body[gi] - 1);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j2 = 0; j2 < 3; ++j2) Re[3 * i + j2] = Rc[i] * Rb[j2] + Rc[3 + i] * Rb[3 + j2] + Rc[6 + i] * Rb[6 + j2];
      } else {
#pragma unroll
        for (int i = 0; i < 9; ++i) Re[i] = (i % 4 == 0) ? 1. : 0.;
      }
#pragma unroll
      for (int i = 0; i < 9; ++i) { sRe[9 * gi + i] = Re[i]; FT(b.fRe, 9 * gi + i) = Re[i]; }
    }
    // constraints, from the last lane downwards (the first lanes hold two bodies)
    for (int c = ARB_GL - 1 - lane; c < m.nc; c += ARB_GL) {
      const int* ci = m.cint + ARB_CONS_NINT * c;
      const int type = m.ctype[c];
      const int r0 = m.crow[c];
      FT(b.factive, c) = 0;
      FT(b.fbranch, c) = 0;
      if (!ci[3]) continue;
      double aux[4] = {0., 0., 0., 0.}, T1[24], T0[24];
      int zi[3] = {0, 0, 0};
      bool act;
      if (type == ARB_CONS_JOINT_LIMITS) {
        Se3 I;
        se3_identity(I);
        act = constraint_update(m, c, I, I, nullptr, nullptr, ST_LD(b.gpos, ci[2]), dt, aux, T1, T0, zi);
        ST(b.cforce, r0) = 0.;
      } else {
        Se3 P0, P1;
        double TW0[6], TW1[6];
        if (ci[0] == 0) se3_identity(P0); else grp_load_se3(sPose + 12 * (ci[0] - 1), P0);
        if (ci[1] == 0) se3_identity(P1); else grp_load_se3(sPose + 12 * (ci[1] - 1), P1);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          TW0[i] = (ci[0] == 0) ? 0. : sT[6 * (ci[0] - 1) + i];
          TW1[i] = (ci[1] == 0) ? 0. : sT[6 * (ci[1] - 1) + i];
        }
        const bool aligned = m.caligned[c] != 0;
        act = constraint_update(m, c, P0, P1, TW0, TW1, 0., dt, aux, T1, T0, zi, aligned);
        if (type == ARB_CONS_SOFT_FINGER_PLANE_POINT) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ST(b.cforce, r0 + i) = 0.;
#pragma unroll
          for (int i = 0; i < 3; ++i) FT(b.fzidx, 3 * c + i) = zi[i];
        }
        if (act) {
          if (aligned) {
#pragma unroll
            for (int i = 0; i < 3; ++i) FT(b.fT1, c * 24 + i) = T1[i];
          } else {
            const int nr = arb_cons_ndol(type);
            for (int i = 0; i < nr * 6; ++i) { FT(b.fT1, c * 24 + i) = T1[i]; FT(b.fT0, c * 24 + i) = T0[i]; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) FT(b.faux, 4 * c + i) = aux[i];
      FT(b.factive, c) = act ? 1 : 0;
      if (act) sFlag[0] = 1.;
    }
  GRP_SYNC(g)

  // ---- factorisation, leaf to root (artic_factor) ------------------------------------------------
  // lanes 0-5: row r and column r of IA (and beta_r); lanes 8-13: row r and column r of IM
  GRP_DECLA(double, row, 6);
  GRP_DECLA(double, col, 6);
  GRP_DECL(double, betar);
  bool carry = false;
  int buf = 0;
  for (int j = nj - 1; j >= 0; --j) {
    const int type = m.jtype[j];
    const int par = m.jparent[j];
    const int nd = arb_joint_ndof(type);
    const int dof = m.jdof[j];
    const int flags = m.bflags[j];
    for (int c = nd - 1; c >= 0; --c) {
      const int k = dof + c;
      GRP_LANES(g)
        const int r = lane & 7;
        const bool isM = (lane >> 3) != 0;
        if (r < 6) {
          if (c == nd - 1) {      // start of the joint: body term and children's contributions
            if (!carry) {
#pragma unroll
              for (int i = 0; i < 6; ++i) { GRP_V(row)[i] = 0.; GRP_V(col)[i] = 0.; }
              GRP_V(betar) = 0.;
            }
            if (!isM) {
#pragma unroll
              for (int i = 0; i < 6; ++i) { GRP_V(row)[i] += sAb[42 * j + 6 * r + i]; GRP_V(col)[i] += sAb[42 * j + 6 * i + r]; }
              GRP_V(betar) += sAb[42 * j + 36 + r];
            } else if (flags & ARB_BODY_HASMASS) {
              const double* Mb = m.bmass + 36 * j;
#pragma unroll
              for (int i = 0; i < 6; ++i) { GRP_V(row)[i] += Mb[6 * r + i]; GRP_V(col)[i] += Mb[6 * i + r]; }
            }
            for (int ch = m.jchild0[j]; ch >= 0; ch = m.jsib[ch]) {
              if (carry && ch == j + 1) continue;
              const double* ps = sSlot + 78 * m.gslot[ch] + (isM ? 36 : 0);
#pragma unroll
              for (int i = 0; i < 6; ++i) { GRP_V(row)[i] += ps[6 * r + i]; GRP_V(col)[i] += ps[6 * i + r]; }
              if (!isM) GRP_V(betar) += sSlot[78 * m.gslot[ch] + 72 + r];
            }
          }
          double s[6], v[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) { s[i] = sS[6 * k + i]; v[i] = isM ? sSh[6 * k + i] : s[i]; }
          double up = 0., lu = 0.;
#pragma unroll
          for (int i = 0; i < 6; ++i) { up += GRP_V(row)[i] * v[i]; lu += s[i] * GRP_V(col)[i]; }
          sUx[12 * buf + (isM ? 6 : 0) + r] = up;
          sLx[12 * buf + (isM ? 6 : 0) + r] = lu;
          if (!isM) sBx[6 * buf + r] = GRP_V(betar);
        }
      GRP_SYNC(g)
      GRP_LANES(g)
        const int r = lane & 7;
        const bool isM = (lane >> 3) != 0;
        if (r < 6) {
          double s[6], U[6], Lv[6];
          double d = 0., sb = 0.;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            s[i] = sS[6 * k + i];
            U[i] = sUx[12 * buf + i] + sUx[12 * buf + 6 + i];
            d += s[i] * U[i];
            sb += s[i] * sBx[6 * buf + i];
          }
          if (m.has_pd) d += artic_pd_diag(m, b, w, k, dt);
          if (lane == 0 && !(fabs(d) > 0.)) sFlag[1] = 1.;
          const double dinv = 1. / d;
#pragma unroll
          for (int i = 0; i < 6; ++i) Lv[i] = sLx[12 * buf + (isM ? 6 : 0) + i] * dinv;
          const double u = (artic_tau(m, b, w, k) - sb) * dinv;
          if (!isM) { sU[6 * k + r] = U[r]; sLA[6 * k + r] = Lv[r]; }
          else sLM[6 * k + r] = Lv[r];
          if (lane == 0) { su0[k] = u; sdinv[k] = dinv; }
          if (c > 0 || par != 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { GRP_V(row)[i] -= U[r] * Lv[i]; GRP_V(col)[i] -= U[i] * Lv[r]; }
            if (!isM) GRP_V(betar) += U[r] * u;
          }
        }
      GRP_END(g)   // (no barrier: the next exchange uses the other buffer)
      buf ^= 1;
    }
    if (par != 0) {
      // parent: X^T IA X, X^T IM X, X^T beta.  B = A X from the rows, C = X^T A from the columns,
      // exchanged through shared memory: the new column r is X^T col_r(B), the new row r is row_r(C) X
      GRP_LANES(g)
        const int r = lane & 7;
        const bool isM = (lane >> 3) != 0;
        if (r < 6) {
          Se3 X;
          grp_load_se3(sX + 12 * j, X);
          double y[6], z[6];
          wrench_up(X, GRP_V(row), y);
          wrench_up(X, GRP_V(col), z);
          double* pb = sBm + (isM ? 36 : 0);
          double* pc = sCm + (isM ? 36 : 0);
#pragma unroll
          for (int i = 0; i < 6; ++i) { pb[6 * r + i] = y[i]; pc[6 * i + r] = z[i]; }
          if (!isM) sBx[6 * buf + r] = GRP_V(betar);
        }
      GRP_SYNC(g)
      const bool keep = (par == j);       // the parent body is the one of joint j-1: next iteration
      GRP_LANES(g)
        const int r = lane & 7;
        const bool isM = (lane >> 3) != 0;
        if (r < 6) {
          Se3 X;
          grp_load_se3(sX + 12 * j, X);
          const double* pb = sBm + (isM ? 36 : 0);
          const double* pc = sCm + (isM ? 36 : 0);
          double bc[6], cr[6], ncol[6], nrow[6];
#pragma unroll
          for (int i = 0; i < 6; ++i) { bc[i] = pb[6 * i + r]; cr[i] = pc[6 * r + i]; }
          wrench_up(X, bc, ncol);
          wrench_up(X, cr, nrow);
#pragma unroll
          for (int i = 0; i < 6; ++i) { GRP_V(row)[i] = nrow[i]; GRP_V(col)[i] = ncol[i]; }
          if (!isM) {
            double be[6], bu[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) be[i] = sBx[6 * buf + i];
            wrench_up(X, be, bu);
            GRP_V(betar) = bu[r];
          }
          if (!keep) {
            double* ps = sSlot + 78 * m.gslot[j] + (isM ? 36 : 0);
#pragma unroll
            for (int i = 0; i < 6; ++i) ps[6 * r + i] = GRP_V(row)[i];
            if (!isM) sSlot[78 * m.gslot[j] + 72 + r] = GRP_V(betar);
          }
        }
      GRP_SYNC(g)
      carry = keep;
      buf ^= 1;
    } else {
      carry = false;
      GRP_LANES(g)
      GRP_SYNC(g)          // the outputs of this joint's last dof are read by the solves
    }
  }

  // ---- solves: one lane per right-hand side ------------------------------------------------------
  // column 0: the free motion (u of the factorisation); columns 1..NG: the generators
  const bool any = sFlag[0] != 0.;
  const int ncol = any ? NG + 1 : 1;
  const int VS = ARB_GL * 12;
  for (int cb = 0; cb < ncol; cb += ARB_GL) {
    GRP_LANES(g)
      const int col_ = cb + lane;
      if (col_ < ncol) {
        const int gc = col_ - 1;                    // generator row (-1: free motion)
        int off = 0, l0 = -1, gi = -1;
        double* au = sAu + lane * L.maxpath;
        if (gc >= 0) {
          // leaf-to-root along the generator's root path (artic_backward_generators)
          double beta[6] = {0., 0., 0., 0., 0., 0.};
          int kstart = -1;
          if (gc < 6 * m.ngen) {
            gi = gc / 6;
            const int r = gc % 6;
            const int body = m.gen_body[gi];
            off = m.coloff[body];
            l0 = m.kcols[body] - 1;
            if (m.gen_aligned[gi]) {
#pragma unroll
              for (int i = 0; i < 6; ++i) beta[i] = ((r < 3) == (i < 3)) ? -sRe[9 * gi + 3 * (r % 3) + (i % 3)] : 0.;
            } else {
#pragma unroll
              for (int i = 0; i < 6; ++i) beta[i] = (i == r) ? -1. : 0.;
            }
          } else {
            kstart = m.glimdof[gc - 6 * m.ngen];
            off = m.coloff[m.dofbody[kstart]];
            l0 = m.dofpos[kstart];
          }
          for (int l = l0; l >= 0; --l) {
            const int k = m.pathdof[off + l];
            const int j = m.dofjoint[k];
            double sb = 0.;
#pragma unroll
            for (int i = 0; i < 6; ++i) sb += sS[6 * k + i] * beta[i];
            const double tau = (k == kstart) ? 1. : 0.;
            const double u = (tau - sb) * sdinv[k];
            au[l] = u;
            if (l > 0) {
#pragma unroll
              for (int i = 0; i < 6; ++i) beta[i] += sU[6 * k + i] * u;
            }
            if (k == m.jdof[j] && m.jparent[j] != 0) {
              Se3 X;
              grp_load_se3(sX + 12 * j, X);
              double y[6];
              wrench_up(X, beta, y);
#pragma unroll
              for (int i = 0; i < 6; ++i) beta[i] = y[i];
            }
          }
        }
        // root-to-leaf over ALL joints (artic_forward_full): x_k = u_k - LA_k V - LM_k V^
        double V[6], Vh[6];
        int prev = -2;
        for (int j = 0; j < nj; ++j) {
          const int par = m.jparent[j];
          const int nd = arb_joint_ndof(m.jtype[j]);
          const int dof = m.jdof[j];
          if (par == 0) {
#pragma unroll
            for (int i = 0; i < 6; ++i) { V[i] = 0.; Vh[i] = 0.; }
          } else {
            Se3 X;
            grp_load_se3(sX + 12 * j, X);
            if (par - 1 != prev) {
              const double* pv = sAv + m.gvslot[par - 1] * VS + lane * 12;
#pragma unroll
              for (int i = 0; i < 6; ++i) { V[i] = pv[i]; Vh[i] = pv[6 + i]; }
            }
            artic_down(X, V, Vh);
          }
          for (int c = 0; c < nd; ++c) {
            const int k = dof + c;
            double t;
            if (gc < 0) {
              t = su0[k];
            } else {
              const int pos = m.dofpos[k];
              t = (pos <= l0 && m.pathdof[off + pos] == k) ? au[pos] : 0.;
            }
            if (par != 0 || c > 0) {
#pragma unroll
              for (int i = 0; i < 6; ++i) t -= sLA[6 * k + i] * V[i] + sLM[6 * k + i] * Vh[i];
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) { V[i] += sS[6 * k + i] * t; Vh[i] += sSh[6 * k + i] * t; }
            if (gc < 0) FT(b.fq, k) = t; else FT(b.fK, k * NG + gc) = t;
            if (any) {      // joint-limit generator rows read the dof itself
              for (int h = 6 * m.ngen; h < NG; ++h)
                if (m.glimdof[h - 6 * m.ngen] == k) {
                  if (gc < 0) FT(b.fv0, h) = t; else FT(b.fLam, h * NG + gc) = t;
                }
            }
          }
          if (m.gvslot[j] >= 0) {
            double* pv = sAv + m.gvslot[j] * VS + lane * 12;
#pragma unroll
            for (int i = 0; i < 6; ++i) { pv[i] = V[i]; pv[6 + i] = Vh[i]; }
          }
          if (any && m.jmark[j]) {      // generator rows of this body: its twist, in the frame R_e when aligned
            for (int g2 = 0; g2 < m.ngen; ++g2)
              if (m.gen_body[g2] == j + 1) {
                double out[6];
                if (m.gen_aligned[g2]) {
                  m3_mulv(sRe + 9 * g2, V, out);
                  m3_mulv(sRe + 9 * g2, V + 3, out + 3);
                } else {
#pragma unroll
                  for (int i = 0; i < 6; ++i) out[i] = V[i];
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                  if (gc < 0) FT(b.fv0, 6 * g2 + i) = out[i]; else FT(b.fLam, (6 * g2 + i) * NG + gc) = out[i];
                }
              }
          }
          prev = j;
        }
      }
    GRP_SYNC(g)
  }
  GRP_LANES(g)
    if (lane == 0 && sFlag[1] != 0.) b.status[w] |= ARB_STATUS_SINGULAR;
  GRP_SYNC(g)
}

// ---------------------------------------------------------------------------------------
// finish stage after the group prepare stage: q'+ = q_free + K y (K = Z^-1 G^T), then the joint
// integration of world_fused_finish (core.py:974-980).  One world per lane.
ARB_D void world_fused_finish_k(const DevModel& m, const DevBatch& b, int64_t w, double dt) {
  const int NG = m.ngrows;
  bool any = false;
  for (int c = 0; c < m.nc; ++c) any = any || FT(b.factive, c);
  double y[32];
  bool useK = any && NG <= 32;
  if (useK) {
    bool nz = false;
    for (int g = 0; g < NG; ++g) { y[g] = FT(b.fy, g); nz = nz || (y[g] != 0.); }
    useK = nz;
  }
  bool finite = true;
  for (int j = 0; j < m.nj; ++j) {
    const int type = m.jtype[j];
    const int gp = m.jgpos[j], d = m.jdof[j];
    const int nd = arb_joint_ndof(type);
    double nv[6];
    for (int i = 0; i < nd; ++i) {
      double t = FT(b.fq, d + i);
      if (useK) {
        double acc = 0.;
        for (int g = 0; g < NG; ++g) acc += FT(b.fK, (d + i) * NG + g) * y[g];
        t += acc;
      } else if (any && NG > 32) {
        double acc = 0.;
        for (int g = 0; g < NG; ++g) acc += FT(b.fK, (d + i) * NG + g) * FT(b.fy, g);
        t += acc;
      }
      ST(b.gvel, d + i) = t;
      finite = finite && isfinite(t);
      nv[i] = t;
    }
    if (type == ARB_JOINT_FREE) {
      double q[16], tw[6];
      for (int i = 0; i < 16; ++i) q[i] = ST_LD(b.gpos, gp + i);
#pragma unroll
      for (int i = 0; i < 6; ++i) tw[i] = dt * nv[i];
      Se3 H, E, R;
      se3_from16(q, H);
      se3_exp(tw, E);
      se3_mul(H, E, R);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) ST(b.gpos, gp + 4 * r + c) = R.R[3 * r + c];
        ST(b.gpos, gp + 4 * r + 3) = R.p[r];
      }
    } else {
      for (int i = 0; i < nd; ++i) ST(b.gpos, gp + i) = ST_LD(b.gpos, gp + i) + dt * nv[i];
    }
  }
  if (!finite) b.status[w] |= ARB_STATUS_NONFINITE;
}
