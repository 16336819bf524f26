/* b200sim.h - C ABI of the B200-native world-batched physics step.
 *
 * Drop-in boundary for the five calls (and two attribute bags) through which
 * mjlab reaches mujoco_warp (reference: src/mjlab/sim/sim.py):
 *
 *   mjwarp.put_model(mjm)                      sim.py:139      -> b2_model (filled by the host shim)
 *   mjwarp.put_data(mjm, mjd, nworld, ...)     sim.py:143-149  -> b2_data  (filled by the host shim)
 *   mjwarp.step(m, d)                          sim.py:184,273  -> b2_step
 *   mjwarp.forward(m, d)                       sim.py:187,265  -> b2_forward
 *   mjwarp.reset_data(m, d, reset=mask)        sim.py:190,286  -> b2_reset
 *   repeat_array_kernel (expand_model_fields)  sim/randomization.py:9-49 -> b2_repeat_rows
 *
 * Conventions: plain pointers and sizes only, no torch types.
 * All device buffers are owned by the caller (the Python shim allocates them
 * as torch tensors and passes raw pointers), stay at stable addresses, and are
 * world-major: field f of world w starts at ptr + w*row_len.  Every call
 * enqueues work on the caller's CUDA stream, performs no allocation and no
 * synchronisation, and touches no host-visible state, so it can be captured
 * into a CUDA graph (sim.py:164-191).
 *
 * Model float fields carry a world stride: 0 = one row shared by all worlds
 * (the reference's leading-dim-1/stride-0 convention, sim_data.py:23-31),
 * row_len = per-world rows after expand_model_fields.
 */
#ifndef B200SIM_H_
#define B200SIM_H_

#ifdef __cplusplus
extern "C" {
#endif

/* ---- model float fields (row length per world in comments) ------------- */
enum b2_mf {
  B2_MF_OPT_TIMESTEP = 0, /* 1 */
  B2_MF_OPT_GRAVITY,      /* 3 */
  B2_MF_OPT_TOLERANCE,    /* 1 */
  B2_MF_OPT_LS_TOLERANCE, /* 1 */
  B2_MF_OPT_IMPRATIO,     /* 1 */
  B2_MF_STAT_MEANINERTIA, /* 1 */
  B2_MF_QPOS0,            /* nq */
  B2_MF_QPOS_SPRING,      /* nq */
  B2_MF_BODY_POS,         /* nbody*3 */
  B2_MF_BODY_QUAT,        /* nbody*4 */
  B2_MF_BODY_IPOS,        /* nbody*3 */
  B2_MF_BODY_IQUAT,       /* nbody*4 */
  B2_MF_BODY_MASS,        /* nbody */
  B2_MF_BODY_SUBTREEMASS, /* nbody */
  B2_MF_BODY_INERTIA,     /* nbody*3 */
  B2_MF_BODY_INVWEIGHT0,  /* nbody*2 */
  B2_MF_JNT_POS,          /* njnt*3 */
  B2_MF_JNT_AXIS,         /* njnt*3 */
  B2_MF_JNT_RANGE,        /* njnt*2 */
  B2_MF_JNT_SOLREF,       /* njnt*2 */
  B2_MF_JNT_SOLIMP,       /* njnt*5 */
  B2_MF_JNT_MARGIN,       /* njnt */
  B2_MF_JNT_STIFFNESS,    /* njnt */
  B2_MF_DOF_ARMATURE,     /* nv */
  B2_MF_DOF_DAMPING,      /* nv */
  B2_MF_DOF_INVWEIGHT0,   /* nv */
  B2_MF_GEOM_POS,         /* ngeom*3 */
  B2_MF_GEOM_QUAT,        /* ngeom*4 */
  B2_MF_GEOM_SIZE,        /* ngeom*3 */
  B2_MF_GEOM_RBOUND,      /* ngeom */
  B2_MF_GEOM_FRICTION,    /* ngeom*3 */
  B2_MF_GEOM_SOLREF,      /* ngeom*2 */
  B2_MF_GEOM_SOLIMP,      /* ngeom*5 */
  B2_MF_GEOM_SOLMIX,      /* ngeom */
  B2_MF_GEOM_MARGIN,      /* ngeom */
  B2_MF_GEOM_GAP,         /* ngeom */
  B2_MF_SITE_POS,         /* nsite*3 */
  B2_MF_SITE_QUAT,        /* nsite*4 */
  B2_MF_ACTUATOR_GAINPRM, /* nu*10 */
  B2_MF_ACTUATOR_BIASPRM, /* nu*10 */
  B2_MF_ACTUATOR_GEAR,    /* nu*6 */
  B2_MF_ACTUATOR_FORCERANGE, /* nu*2 */
  B2_MF_ACTUATOR_CTRLRANGE,  /* nu*2 */
  B2_MF_SENSOR_CUTOFF,    /* nsensor */
  B2_MF_SGEOM_XPOS,       /* ngeom*3: world position of static geoms (bodies welded to the world); 0 for moving geoms */
  B2_MF_SGEOM_XMAT,       /* ngeom*9 */
  B2_MF_FIELD_AABB,       /* nfield*6: world AABB (centre, half extents) of each field geom */
  B2_MF_FIELD_GRID,       /* 4: x0, y0, cell edge, 1 / cell edge of the uniform xy grid over the field */
  B2_MF_DOF_FRICTIONLOSS, /* nv: dry-friction force bound per dof (> 0 gives the dof a friction-loss constraint row) */
  B2_MF_DOF_SOLREF,       /* nv*2: solref of the friction-loss row (MuJoCo's dof_solref, <joint solreffriction>) */
  B2_MF_DOF_SOLIMP,       /* nv*5: solimp of the friction-loss row (dof_solimp, <joint solimpfriction>) */
  B2_MF_COUNT
};

/* ---- model int fields (shared by all worlds) ---------------------------- */
enum b2_mi {
  B2_MI_BODY_PARENTID = 0, /* nbody */
  B2_MI_BODY_ROOTID,       /* nbody */
  B2_MI_BODY_MOCAPID,      /* nbody */
  B2_MI_BODY_JNTNUM,       /* nbody */
  B2_MI_BODY_JNTADR,       /* nbody */
  B2_MI_BODY_DOFNUM,       /* nbody */
  B2_MI_BODY_DOFADR,       /* nbody */
  B2_MI_BODY_LASTDOF,      /* nbody: last dof on the path to the root, -1 if none */
  B2_MI_BODY_DOFMASK,      /* nbody*2: 64-bit mask of ancestor dofs (lo, hi) */
  B2_MI_LEVEL_ADR,         /* nlevel+1: CSR over tree levels */
  B2_MI_LEVEL_BODY,        /* nbody: body ids sorted by level */
  B2_MI_CHILD_ADR,         /* nbody+1: CSR over children */
  B2_MI_CHILD_BODY,        /* nbody-1 */
  B2_MI_JNT_TYPE,          /* njnt */
  B2_MI_JNT_QPOSADR,       /* njnt */
  B2_MI_JNT_DOFADR,        /* njnt */
  B2_MI_JNT_BODYID,        /* njnt */
  B2_MI_JNT_LIMITED,       /* njnt */
  B2_MI_DOF_BODYID,        /* nv */
  B2_MI_DOF_JNTID,         /* nv */
  B2_MI_DOF_PARENTID,      /* nv */
  B2_MI_GEOM_TYPE,         /* ngeom */
  B2_MI_GEOM_BODYID,       /* ngeom */
  B2_MI_GEOM_CONDIM,       /* ngeom */
  B2_MI_GEOM_PRIORITY,     /* ngeom */
  B2_MI_SITE_BODYID,       /* nsite */
  B2_MI_PAIR_GEOM,         /* npair*2: static candidate pairs, geom1 has the lower type */
  B2_MI_ACTUATOR_TRNID,    /* nu*2 */
  B2_MI_ACTUATOR_BIASTYPE, /* nu */
  B2_MI_ACTUATOR_FORCELIMITED, /* nu */
  B2_MI_ACTUATOR_CTRLLIMITED,  /* nu */
  B2_MI_SENSOR_TYPE,       /* nsensor */
  B2_MI_SENSOR_OBJTYPE,    /* nsensor */
  B2_MI_SENSOR_OBJID,      /* nsensor */
  B2_MI_SENSOR_REFTYPE,    /* nsensor */
  B2_MI_SENSOR_REFID,      /* nsensor */
  B2_MI_SENSOR_INTPRM,     /* nsensor*3 */
  B2_MI_SENSOR_DIM,        /* nsensor */
  B2_MI_SENSOR_ADR,        /* nsensor */
  B2_MI_DOF_ACTID,         /* nv: the actuator driving this dof, -1 none, -2 several (kernels then scan) */
  B2_MI_BODY_ANCMASK,      /* nbody*2: 64-bit mask (lo, hi) of the body itself and its ancestors; valid when nbody <= 64
                              (subtree tests of the contact sensors become one load instead of a walk up the tree) */
  B2_MI_DYN_GEOM,          /* ndyn: geoms that move (forward kinematics runs for these only) */
  B2_MI_FIELD_GEOM,        /* nfield: geom id of each field geom (static terrain boxes, see b2_model.nfield) */
  B2_MI_FIELD_CELL0,       /* nfield*2: lowest grid cell (ix, iy) the geom's AABB touches */
  B2_MI_FIELD_CELL_ADR,    /* field_nx*field_ny+1: CSR offsets into FIELD_CELL_ITEM, cell = ix*field_ny + iy */
  B2_MI_FIELD_CELL_ITEM,   /* field indices whose AABB overlaps the cell, ascending */
  B2_MI_FQ_GEOM,           /* nfq: moving geoms (spheres, capsules) that can touch the field, in geom order */
  B2_MI_COUNT
};

/* ---- data float fields (all [nworld, row]) ------------------------------- */
enum b2_df {
  B2_DF_QPOS = 0,        /* nq */
  B2_DF_QVEL,            /* nv */
  B2_DF_CTRL,            /* nu */
  B2_DF_QACC,            /* nv */
  B2_DF_QACC_WARMSTART,  /* nv */
  B2_DF_QFRC_APPLIED,    /* nv */
  B2_DF_XFRC_APPLIED,    /* nbody*6 */
  B2_DF_MOCAP_POS,       /* nmocap*3 */
  B2_DF_MOCAP_QUAT,      /* nmocap*4 */
  B2_DF_TIME,            /* 1 */
  B2_DF_XPOS,            /* nbody*3 */
  B2_DF_XQUAT,           /* nbody*4 */
  B2_DF_XMAT,            /* nbody*9 */
  B2_DF_XIPOS,           /* nbody*3 */
  B2_DF_XIMAT,           /* nbody*9 */
  B2_DF_XANCHOR,         /* njnt*3 */
  B2_DF_XAXIS,           /* njnt*3 */
  B2_DF_GEOM_XPOS,       /* ngeom*3 */
  B2_DF_GEOM_XMAT,       /* ngeom*9 */
  B2_DF_SITE_XPOS,       /* nsite*3 */
  B2_DF_SITE_XMAT,       /* nsite*9 */
  B2_DF_SUBTREE_COM,     /* nbody*3 */
  B2_DF_CINERT,          /* nbody*10 */
  B2_DF_CDOF,            /* nv*6 */
  B2_DF_CDOF_DOT,        /* nv*6 */
  B2_DF_CVEL,            /* nbody*6 */
  B2_DF_CACC,            /* nbody*6 */
  B2_DF_QM,              /* nv*ld dense, row stride ld = b2 leading dimension (>= nv, ld % 8 == 4) */
  B2_DF_QFRC_BIAS,       /* nv */
  B2_DF_QFRC_PASSIVE,    /* nv */
  B2_DF_QFRC_ACTUATOR,   /* nv */
  B2_DF_QFRC_SMOOTH,     /* nv */
  B2_DF_QACC_SMOOTH,     /* nv */
  B2_DF_QFRC_CONSTRAINT, /* nv */
  B2_DF_ACTUATOR_FORCE,  /* nu */
  B2_DF_ACTUATOR_LENGTH, /* nu */
  B2_DF_ACTUATOR_VELOCITY, /* nu */
  B2_DF_SENSORDATA,      /* nsensordata */
  B2_DF_SUBTREE_LINVEL,  /* nbody*3 */
  B2_DF_SUBTREE_ANGMOM,  /* nbody*3 */
  B2_DF_CONTACT_DIST,    /* nconmax */
  B2_DF_CONTACT_POS,     /* nconmax*3 */
  B2_DF_CONTACT_FRAME,   /* nconmax*9 */
  B2_DF_CONTACT_INCLUDEMARGIN, /* nconmax */
  B2_DF_CONTACT_FRICTION, /* nconmax*5 */
  B2_DF_CONTACT_SOLREF,  /* nconmax*2 */
  B2_DF_CONTACT_SOLIMP,  /* nconmax*5 */
  B2_DF_EFC_J,           /* njmax*nv */
  B2_DF_EFC_D,           /* njmax */
  B2_DF_EFC_AREF,        /* njmax */
  B2_DF_EFC_POS,         /* njmax */
  B2_DF_EFC_MARGIN,      /* njmax */
  B2_DF_EFC_FORCE,       /* njmax */
  B2_DF_MEMO_QPOS,       /* nq: the qpos the position / velocity stage outputs of this world were computed from; internal */
  B2_DF_MEMO_QVEL,       /* nv: likewise for qvel; internal */
  B2_DF_COUNT
};

/* ---- data int fields (all [nworld, row]) --------------------------------- */
enum b2_di {
  B2_DI_NCON = 0,          /* 1 */
  B2_DI_NEFC,              /* 1 */
  B2_DI_NL,                /* 1: number of joint-limit rows (efc rows nf..nf+nl-1, after the friction-loss rows) */
  B2_DI_SOLVER_NITER,      /* 1 */
  B2_DI_CONTACT_GEOM,      /* nconmax*2 */
  B2_DI_CONTACT_DIM,       /* nconmax */
  B2_DI_CONTACT_EFC_ADDRESS, /* nconmax */
  B2_DI_NCON_DROPPED,      /* 1: contacts dropped because nconmax was reached */
  B2_DI_NEFC_DROPPED,      /* 1: rows dropped because njmax was reached */
  B2_DI_SOLVE_ORDER,       /* 1: world processed by solver slot w (worlds sorted by contact count, heaviest first, so the
                              long-running worlds start at once and do not stretch the kernel's tail); internal */
  B2_DI_SOLVE_NEXT,        /* 1: element [w0] is the work counter of the solver launch over worlds [w0, w1): its warps pull
                              solver slots from it until the list is empty (set by the ordering kernel); internal */
  B2_DI_SOLVE_RING,        /* 1: elements [w0, w1): the worlds of that launch in the order their solve finished (-1 = not yet);
                              warps that find the solver queue empty take the sensor / integration stage of finished worlds
                              from here instead of idling through the launch's tail; internal */
  B2_DI_SOLVE_CTR,         /* 4: row w0 = {worlds solved, next ring entry to post-process, -, -} of that launch; internal */
  B2_DI_SMOOTH_MEMO,       /* 1: 1 = MEMO_QPOS / MEMO_QVEL describe the stage outputs in HBM; write 0 after changing MODEL fields; internal */
  B2_DI_NF,                /* 1: number of friction-loss rows (efc rows 0..nf-1; MuJoCo's order: friction, limits, contacts) */
  B2_DI_COUNT
};

typedef struct b2_model {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, nsensor, nsensordata, npair, nmocap, nlevel;
  int integrator;    /* 0 euler, 3 implicitfast */
  int cone;          /* 0 pyramidal, 1 elliptic (MujocoCfg.cone, src/mjlab/sim/sim.py:51-56) */
  int solver;        /* 2 newton (only supported solver) */
  int iterations, ls_iterations, disableflags;
  int has_accelerometer, has_subtreevel;
  int nlimited;      /* number of limited hinge/slide joints (bounds the joint-limit rows) */
  int maxchain;      /* max over candidate pairs of |dofs(body1) U dofs(body2)| (compact row width bound) */
  int nfriction;     /* capacity for friction-loss rows: 0 = the model has none (the solver kernels without the friction
                      * code are launched), else an upper bound on the dofs with dof_frictionloss > 0 in any world */
  /* geom field: a large set of static boxes (the rough-terrain grid, terrains/terrain_generator.py) is not enumerated as
   * candidate pairs; moving "query" geoms look the boxes up in a uniform xy grid.  nfield == 0: no field. */
  int nfield, nfq, ndyn, field_nx, field_ny;
  const float* mf[B2_MF_COUNT];
  int mf_stride[B2_MF_COUNT]; /* floats between consecutive worlds; 0 = shared */
  const int* mi[B2_MI_COUNT];
} b2_model;

typedef struct b2_data {
  int nworld, nconmax, njmax;
  int flags;         /* bit 0: also write the dense efc_J rows to HBM (tests / API consumers); off on the hot path */
  float* df[B2_DF_COUNT];
  int* di[B2_DI_COUNT];
} b2_data;

/* Library / build identification. */
int b2_abi_version(void);
const char* b2_last_error(void);
/* sizeof checks for the FFI side */
int b2_sizeof_model(void);
int b2_sizeof_data(void);
/* solver slot of the world with heaviest-first rank `rank` in a launch over `nworld` worlds (diagnostic: the scheduling
 * map of k_order, exposed so that it can be checked on a machine without a GPU) */
int b2_debug_solver_slot(int rank, int nworld);
/* number of kernel launches enqueued by this library since load (diagnostic) */
long long b2_launch_count(void);

/* Validates sizes against compiled-in limits and the device's shared-memory
 * capacity; returns 0 on success, non-zero + b2_last_error() otherwise.
 * (mirrors the setup-time ValueError behaviour, sim.py:231-233) */
int b2_check(const b2_model* m, const b2_data* d);

/* mjwarp.step: forward dynamics + constraint solve + sensors + integration.
 * mjwarp.forward: the same without integration (time, qpos, qvel, warm start untouched).
 * stream is a cudaStream_t passed as void*.  Return 0 on successful enqueue. */
int b2_step(const b2_model* m, b2_data* d, void* stream);
int b2_forward(const b2_model* m, b2_data* d, void* stream);

/* nsub consecutive b2_step calls (the decimation loop of an env step) in one call.  With B200SIM_CHUNKS=N (default 1) the
 * batch is cut into N chunks whose kernel chains run on library-owned side streams forked from / joined to `stream` (graph
 * edges under stream capture); results are identical for any N. */
int b2_substeps(const b2_model* m, b2_data* d, int nsub, void* stream);

/* b2_substeps for the caller that knows a b2_forward came last (ManagerBasedRlEnv.step: the forward pass after the resets,
 * manager_based_rl_env.py:306-311, then the next step's decimation loop, :275-280).  Each world whose qpos / qvel are still
 * bit for bit those of that pass (B2_DF_MEMO_QPOS / _QVEL; B2_DI_SMOOTH_MEMO == 1; no external wrench, no mocap body) skips
 * the position / velocity stage of the FIRST sub-step - kinematics, mass matrix, collision, bias forces are in HBM already -
 * and only redoes what depends on ctrl (actuator forces, qfrc_smooth, qacc_smooth) through the same instructions, so results
 * are bit-identical to b2_substeps.  MODEL fields edited in between are not seen by that comparison: zero
 * B2_DI_SMOOTH_MEMO after such an edit (Simulation.invalidate_memo()).  B200SIM_MEMO=0 turns the reuse off. */
int b2_substeps_after_forward(const b2_model* m, b2_data* d, int nsub, void* stream);

/* One stage of b2_step on its own: 0 = k_smooth (position / velocity stage) + k_order (solver schedule), 1 = k_solve
 * (constraint rows + Newton solver), 2 = k_post (sensors, and integration when do_integrate != 0).  b2_step == stages 0, 1, 2
 * with do_integrate = 1; b2_forward the same with do_integrate = 0.  4 = k_smooth alone, 5 = k_order alone (0 == 4 then 5);
 * 3 = k_solve with the sensor stage fused into its tail (experiment); 6 = k_smooth the way b2_substeps_after_forward launches
 * its first sub-step.  Used by bench.py to bracket each kernel with CUDA
 * events and by stage-wise tests. */
int b2_stage(const b2_model* m, b2_data* d, int stage, int do_integrate, void* stream);

/* mjwarp.reset_data(m, d, reset=mask): worlds with mask[w] != 0 (bool bytes)
 * return to qpos0 / zero velocity, controls, warm start, applied forces, time;
 * mask == NULL resets every world. */
int b2_reset(const b2_model* m, b2_data* d, const unsigned char* mask, void* stream);

/* Writes the world pose of the static geoms (B2_MF_SGEOM_*) into every world's geom_xpos / geom_xmat rows.  The step
 * kernels only update moving geoms, so call this once after the data buffers are allocated. */
int b2_init_static(const b2_model* m, b2_data* d, void* stream);

/* repeat_array_kernel: dst[w, :] = src[:] for w in [0, nworld). */
int b2_repeat_rows(float* dst, const float* src, int row_len, int nworld, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SIM_H_ */
