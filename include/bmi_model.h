/*
 * bmi_model.h — layout of the baked robot model blob (float32[]) passed to bmi_env_create.
 * Produced by tools/bake_model.py from the reference URDF/meshes; consumed by
 * csrc/physics.cu (staged into shared memory with one TMA bulk copy) and by the C oracle
 * (oracle/bmi_physics_oracle.c).  Integer-valued fields are stored as exact floats.
 */
#ifndef BMI_MODEL_H_
#define BMI_MODEL_H_

#define BMI_MODEL_MAGIC 20251017.0f
#define BMI_MODEL_HDR 64
#define BMI_LINK_STRIDE 32
#define BMI_MAX_LINKS 9
#define BMI_SHAPE_STRIDE 12
#define BMI_MAX_SHAPES 4
#define BMI_MODEL_MAX_FLOATS 2048

/* header / parameter slots */
enum {
  MP_MAGIC = 0, MP_VERSION = 1, MP_N_LINKS = 2, MP_N_SHAPES = 3, MP_LINKS_OFF = 4, MP_SHAPES_OFF = 5,
  MP_POOL_OFF = 6, MP_TOTAL = 7,
  MP_DT = 8,              /* 1/240        bmirobot_env_push_F.py:73 */
  MP_GRAVITY = 9,         /* -10          bmirobot_env_push_F.py:161 */
  MP_N_SUBSTEPS = 10,     /* 20           bmirobot_push_F.py:17 */
  MP_SOLVER_ITERS = 11,   /* 150          bmirobot_env_push_F.py:111 */
  MP_RESIDUAL_THRESH = 12,/* PGS early exit on max (delta impulse / invDiag)^2 */
  MP_ERP_JOINT = 13,      /* 0.2  joint-limit ERP */
  MP_ERP_CONTACT = 14,    /* 0.08 contact ERP (pinned by the block-settle transient of the goldens) */
  MP_LINEAR_SLOP = 15,    /* 1e-5 */
  MP_MOTOR_KP = 16,       /* positionGain 0.03   bmirobot.py:160 */
  MP_MOTOR_KD = 17,       /* velocityGain 1      bmirobot.py:161 */
  MP_MOTOR_FORCE = 18,    /* 500                 bmirobot.py:159 */
  MP_LIN_DAMP = 19, MP_ANG_DAMP = 20,   /* Bullet multibody link damping 0.04 */
  MP_IK_DAMPING = 21, MP_IK_ITERS = 22, MP_IK_THRESH = 23, MP_IK_MAX_ANGLE = 24,
  MP_TABLE_Z = 25, MP_MU_TABLE = 26, MP_CONTACT_MARGIN = 27,
  MP_BASE_PX = 28, MP_BASE_PY = 29, MP_BASE_PZ = 30, MP_EE_LINK = 31,
  MP_PUSH_HX = 32, MP_PUSH_HY = 33, MP_PUSH_HZ = 34, MP_PUSH_MASS = 35, MP_PUSH_MU = 36,
  MP_PICK_HX = 37, MP_PICK_HY = 38, MP_PICK_HZ = 39, MP_PICK_MASS = 40, MP_PICK_MU = 41,
  MP_DIST_THRESHOLD = 42, MP_JOINT_LIMIT_IMPULSE = 43, MP_BLOCK_MARGIN = 44, MP_TABLE_MARGIN = 45,
  MP_IK_POS_AT_COM = 46, MP_SELF_COLLISION = 47,
  MP_WARMSTART = 48,      /* contact warm-starting factor (Bullet default 0.85); 0 disables */
  MP_HULL_MARGIN = 49,    /* collision margin of a convex-hull shape (PyBullet URDF meshes: 0.001) */
  MP_SELF_SPLIT_DIAG = 50,/* 1: Bullet's row diagonal for contacts between two links of one multibody (no cross term) */
  MP_SWEEP_ALTERNATE = 51,/* 1: non-contact rows are swept backwards on even solver iterations (Bullet) */
  MP_SELF_NEAR = 53,      /* self-collision pairs closer than this (beyond the margins) still produce a (speculative) row */
  MP_FULL_HULLS = 54,     /* oracle only: arm-table / arm-block contacts from the full hulls of all links (GJK + EPA) */
  MP_SELF_TABLE = 55,     /* 1: self-collision pairs come from the baked pair tables (the kernel's path) instead of GJK / EPA */
  MP_PGS_COMPRESS = 56,   /* K > 1: iteration compression of the under-relaxed rows (kernel schedule; oracle default 0 = plain loop) */
  MP_PGS_TAIL = 57,       /* ... number of plain iterations at the end of the compressed schedule */
  MP_LIMITS_FIRST = 52    /* 1: joint-limit rows precede the motor rows (order the constraints were created in) */
};

/* per-link slots (BMI_LINK_STRIDE floats each) */
enum {
  ML_PARENT = 0,   /* -1 = fixed base (right_link1) */
  ML_JPOS = 1,     /* joint origin in the parent link frame (3) */
  ML_JROT = 4,     /* joint origin rotation, row-major 3x3, parent-from-child at q=0 */
  ML_AXIS = 13,    /* revolute axis in the child frame (3) */
  ML_LO = 16, ML_HI = 17, ML_DAMPING = 18, ML_MASS = 19,
  ML_COM = 20,     /* centre of mass in the link frame (3) */
  ML_INERTIA = 23, /* diagonal inertia about the COM, link axes (3) */
  ML_SHAPE = 26,   /* index of the collision polytope or -1 */
  ML_MU = 27
};

/* per-shape slots (BMI_SHAPE_STRIDE floats each) */
enum {
  MS_LINK = 0, MS_NVERTS = 1, MS_NPLANES = 2, MS_VERT_OFF = 3, MS_PLANE_OFF = 4,
  MS_SPHERE_C = 5, MS_SPHERE_R = 8, MS_MU = 9
};

/* Baked self-collision pair tables (assets/bmirobot_selfcol.bin, float32; tools/bake_selfcol.py).  Every link pair of
 * the arm that can touch is separated by exactly two joints (grandparent / sibling pairs), so its whole narrow phase is
 * a function of two joint angles: the kernel looks it up instead of running GJK / EPA on 300-vertex hulls.
 * Header SC_HDR floats, then SC_DESC floats per pair, then per node 8 floats:
 *   core distance (gap > 0 or -penetration depth; 1e3 = far), unit normal from link B towards link A (3) and witness
 *   point on A (3), both in the frame of link A, pad.   witness on B = xa - n * distance. */
#define BMI_SC_MAGIC 20261017.0f
#define BMI_SC_MAX_PAIRS 4
enum { SC_MAGIC = 0, SC_NPAIRS = 1, SC_TOTAL = 2, SC_HDR = 8, SC_DESC = 16 };
enum {
  SC_LA = 0, SC_LB = 1,    /* links (-1 = right_link1, rigid with the base) */
  SC_JA = 2, SC_JB = 3,    /* the two joints the relative pose depends on */
  SC_A0 = 4, SC_B0 = 5,    /* grid origin */
  SC_H = 6,                /* grid spacing (rad) */
  SC_NA = 7, SC_NB = 8,    /* nodes per axis */
  SC_OFF = 9,              /* float offset of node (0, 0) from the start of the file; node (i, j) at OFF + 8 (i NB + j) */
  SC_MU = 10               /* combined friction of the pair (product, clamped to 10) */
};

/* simulator state vector exposed by bmi_env_get_state / set_state (BMI_ENV_STATE_DIM = 48) */
enum {
  ST_Q = 0,        /* 9 joint angles   */
  ST_QD = 9,       /* 9 joint rates    */
  ST_QT = 18,      /* 9 motor targets  */
  ST_BPOS = 27,    /* block position 3 */
  ST_BQUAT = 30,   /* block orientation x,y,z,w */
  ST_BVEL = 34,    /* block linear velocity 3 */
  ST_BANG = 37,    /* block angular velocity 3 */
  ST_GOAL = 40,    /* goal 3 */
  ST_PAD = 43
};

#endif
