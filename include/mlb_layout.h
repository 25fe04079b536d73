/*
 * mlb_layout.h — layout of the "model block": the flat FP64 array that describes one
 * Monte Carlo case (integrator, environment, rocket stages/components, tables) to the
 * batched trajectory kernels. Built on the host by mapleaf_b200/model.py from a .mapleaf
 * SimDefinition; consumed by libmapleaf_b200.so (CUDA) and by the test oracle.
 *
 * Every entry is a double (integers / enums are stored as exactly-representable doubles).
 * All offsets (…_OFF) are indices into the same array. This header is the single source of
 * truth: the Python side parses the `#define NAME value` lines below.
 *
 * What each field restates in the reference (paths relative to MAPLEAF/):
 *   integrator + adaptive parameters ........ Motion/Integration.py:56-103,144-195,238-329
 *   end condition ........................... SimulationRunners/SingleSimulations.py:205-248
 *   earth / atmosphere / wind / turbulence .. ENV/environment.py:39-103, ENV/{Earth,Atmosphere,MeanWind,Turbulence}Modelling.py
 *   launch rail ............................. ENV/launchRail.py:8-14
 *   stage records ........................... Rocket/stage.py:10-41, Rocket/CompositeObject.py:23-63
 *   component records ....................... Rocket/{noseCone,bodyTube,boatTail,Fins,Propulsion,Recovery,RocketComponents}.py constructors
 */
#ifndef MLB_LAYOUT_H
#define MLB_LAYOUT_H

#define MLB_MAGIC 20261019
#define MLB_VERSION 1

/* ---- header ---------------------------------------------------------------------- */
#define H_MAGIC 0
#define H_VERSION 1
#define H_LEN 2
#define H_INTEGRATOR 3
#define H_DT0 4
#define H_END_TYPE 5
#define H_END_VALUE 6
#define H_MAX_STEPS 7
#define H_AD_TARGET 8
#define H_AD_MINF 9
#define H_AD_MAXF 10
#define H_AD_MINDT 11
#define H_AD_MAXDT 12
#define H_AD_CTRL 13
#define H_AD_P 14
#define H_AD_I 15
#define H_AD_D 16
#define H_AD_SAFETY 17
#define H_EVENT_DT 18
#define H_EARTH 19
#define H_LAT 20
#define H_LON 21
#define H_ELEV 22
#define H_ATM 23
#define H_ATM_CONST 24
#define H_ATM_TAB_OFF 28
#define H_ATM_TAB_N 29
#define H_WIND 30
#define H_WIND_V 31
#define H_HELLMAN_ALPHA 34
#define H_HELLMAN_LIMIT 35
#define H_WIND_TAB_OFF 36
#define H_WIND_TAB_N 37
#define H_TURB 38
#define H_TURB_INTENSITY 39
#define H_TURB_STD 40
#define H_GUST_START 41
#define H_GUST_BLEND 42
#define H_GUST_THICK 43
#define H_GUST_MAG 44
#define H_GUST_DIR 45
#define H_TURB_OFF_CHUTE 48
#define H_RAIL_LEN 49
#define H_RAIL_POS 50
#define H_RAIL_DIR 53
#define H_AREF 56
#define H_FULLY_TURB 57
#define H_MAXDIAM 58
#define H_INIT_STATE 59
#define H_NSTAGES 72
#define H_STAGE_OFF 73
#define H_FINENESS 77
#define H_TABLEAU_OFF 81
#define H_TAB_NSTAGE_ROWS 82
#define H_TAB_NRESULT_ROWS 83
#define H_RECOVERY_OFF 84
#define H_CTRL_OFF 85
#define H_EARTH_ROT 86
#define H_USSTD_OFF 87
#define H_PINK_C0 88
#define H_PINK_C1 89
#define H_PINK_SEED 90
#define H_EARTH_GM 93
#define H_EARTH_RADIUS 94
#define H_ROCKET_ROUGHNESS 95
#define H_ADD_AUTO_BT 96
#define H_PINK_HAS_SEED 97
#define H_NOSE_POS0 98
#define H_FIN_CUBIC_OFF 101
#define H_START_TIME 102
#define H_INERTIA_CONST 103   /* 1: no stage's mass / CG / MOI can change in flight (full overrides or no motor) */
/* WGS84 ellipsoid (EarthModelling.py:229-240): semi-major axis, first eccentricity squared; H_EARTH_RADIUS is the sphere radius of
   the Round model and H_EARTH_GM the gravitational parameter of the selected model */
#define H_WGS_A 104
#define H_WGS_E2 105
/* doubles [0, H_SHARED_LEN) are staged in shared memory by the kernels; N-D interpolation tables of the control system (up to
   hundreds of KB) follow and are read from global memory */
#define H_SHARED_LEN 106
/* Host-derived constants of the device path. Every quotient or cosine below has flight-independent operands; the host evaluates it once
   with the reference's own operation order, the kernels multiply by it (or divide through the stored reciprocal with one correction
   step, mlb_math.cuh divBy) instead of running a full FP64 division per derivative evaluation. The oracle ignores them. */
#define H_RAREF 107           /* 1 / H_AREF */
#define H_FINENESS_F 108      /* 1 + 0.5 / H_FINENESS[k] (AeroFunctions.py:150), k = stages remaining - 1; 4 entries */
#define H_INIT_CG0 112        /* CG of the complete rocket at t = 0 in rocket coordinates: H_INIT_STATE position = H_NOSE_POS0 + rotate(q0, CG)
                                 (rocket.py:251-291); lanes with their own motor factors have their own CG and re-derive it */
#define H_SIZE 128

/* integrator ids (H_INTEGRATOR) */
#define INT_EULER 0
#define INT_RK2_MIDPOINT 1
#define INT_RK2_HEUN 2
#define INT_RK4 3
#define INT_RK4_38 4
#define INT_RK12A 10
#define INT_RK23A 11
#define INT_RK45A 12
#define INT_RK78A 13
#define TABLEAU_STRIDE 14

/* end conditions (H_END_TYPE) */
#define END_APOGEE 0
#define END_ALT_ASCENDING 1
#define END_ALT_DESCENDING 2
#define END_TIME 3

/* adaptive step controller (H_AD_CTRL) */
#define CTRL_PID 0
#define CTRL_ELEMENTARY 1
#define CTRL_CONSTANT 2

/* models */
#define EARTH_NONE 0
#define EARTH_FLAT 1
#define EARTH_ROUND 2
#define EARTH_WGS84 3
#define ATM_CONSTANT 0
#define ATM_USSTD 1
#define ATM_TABULATED 2
#define WIND_CONSTANT 0
#define WIND_HELLMAN 1
#define WIND_INTERPOLATED 2
#define TURB_NONE 0
#define TURB_PINK1D 1
#define TURB_PINK2D 2
#define TURB_PINK3D 3
#define TURB_SINE_GUST 4

/* event types (Rocket/simEventDetector.py:14-19) */
#define EV_NONE 0
#define EV_APOGEE 1
#define EV_ASCENDING_ALT 2
#define EV_DESCENDING_ALT 3
#define EV_MOTOR_BURNOUT 4
#define EV_TIME_REACHED 5
/* event callbacks */
#define CB_RECOVERY_DEPLOY 1
#define CB_STAGE_SEPARATION 2
#define MLB_MAX_EVENTS 6

/* ---- stage record ------------------------------------------------------------------ */
#define S_NCOMP 0
#define S_POS 1
#define S_SEP_TYPE 4
#define S_SEP_VALUE 5
#define S_SEP_DELAY 6
#define S_OVR_FLAGS 7
#define S_OVR_CG 8
#define S_OVR_MASS 11
#define S_OVR_MOI 12
#define S_FIXED_MOI 15
#define S_FIXED_CG 18
#define S_FIXED_MASS 21
#define S_BURN_DURATION 22
#define S_MOTOR 23
#define S_IGNITION0 24
#define S_AUTO_BT_OFF 25
#define S_COMP_OFF 26
#define S_MAX_COMP 30
/* components with time-varying inertia (Motor, TabulatedInertia), as indices into S_COMP_OFF, in the order the reference
   combines them: construction order, not the top-to-bottom order of the force sum (CompositeObject.py:41-50,102-121) */
#define S_NVAR 56
#define S_VAR_IDX 57
#define S_MAX_VAR 4
/* the same two lists in top-to-bottom component order: the reference rebuilds them that way when a stage is dropped
   (Rocket._dropStage -> recomputeFixedMassInertia, rocket.py:476-479, CompositeObject.py:65-77) */
#define S_FIXED2_MOI 61
#define S_FIXED2_CG 64
#define S_FIXED2_MASS 67
#define S_VAR2_IDX 68
#define S_SIZE 72
#define OVR_CG 1
#define OVR_MASS 2
#define OVR_MOI 4

/* ---- component records: common prefix ------------------------------------------------ */
#define C_TYPE 0
#define C_LEN 1
#define C_POS 2
#define C_MOI 5
#define C_CG 8
#define C_MASS 11
#define C_P 12

#define CT_NOSECONE 1
#define CT_BODYTUBE 2
#define CT_FINSET 3
#define CT_MOTOR 4
#define CT_MASS 5
#define CT_RECOVERY 6
#define CT_BOATTAIL 7
#define CT_TRANSITION 8
#define CT_TAB_AERO 9
#define CT_TAB_INERTIA 10
#define CT_JET_DAMPING 11
#define CT_FIXED_FORCE 12
#define CT_AERO_FORCE 13
#define CT_AERO_DAMPING 14

/* *_CRIT_RE / *_ROUGH_CF: roughness-critical Reynolds number and fully-rough skin-friction coefficient of the
   component (AeroFunctions.py:95-104), constant per component, precomputed on the host for the device path */
/* nose cone (Rocket/noseCone.py:58-101) */
#define NC_BASE_AREA 12
#define NC_LENGTH 13
#define NC_WETTED 14
#define NC_CP 15
#define NC_HALF_ANGLE 18
#define NC_SUB 19
#define NC_TRANS 21
#define NC_ROUGH 25
#define NC_CRIT_RE 26
#define NC_ROUGH_CF 27
#define NC_K_WET 28       /* NC_WETTED / Aref (AeroFunctions.py:151) */
#define NC_K_RAD 29       /* (NC_WETTED / NC_LENGTH) / (2 pi), the mean radius of :154 */
#define NC_SIZE 30

/* body tube (Rocket/bodyTube.py:13-33) */
#define BT_LENGTH 12
#define BT_DIAM 13
#define BT_PLANFORM 14
#define BT_WETTED 15
#define BT_CP 16
#define BT_ROUGH 19
#define BT_CRIT_RE 20
#define BT_ROUGH_CF 21
#define BT_K_WET 22
#define BT_K_RAD 23
#define BT_K_PLAN 24      /* BT_PLANFORM / Aref (bodyTube.py:46) */
#define BT_DL 25          /* BT_LENGTH / 25 damping strips (bodyTube.py:71) */
#define BT_STRIP_Z 26     /* z of the 25 damping strip centres in rocket coordinates: position.z + (-dL i - dL / 2) (bodyTube.py:74-77) */
#define BT_SIZE 51

/* boat tail / transition (Rocket/boatTail.py:12-99) */
#define TR_TOP_AREA 12
#define TR_BOTTOM_AREA 13
#define TR_FRONTAL 14
#define TR_LENGTH 15
#define TR_WETTED 16
#define TR_CP 17
#define TR_CD_ADJ 20
#define TR_GROWING 21
#define TR_HALF_ANGLE 22
#define TR_SUB 23
#define TR_TRANS 25
#define TR_ROUGH 29
#define TR_START_D 30
#define TR_END_D 31
#define TR_CRIT_RE 32
#define TR_ROUGH_CF 33
#define TR_K_WET 34
#define TR_K_RAD 35
#define TR_K_FRONT 36     /* TR_FRONTAL / Aref (boatTail.py:153,178) */
#define TR_SIZE 37

/* fin set (Rocket/Fins.py:45-262, 430-462) */
#define FS_NFINS 12
#define FS_SPAN 13
#define FS_ROOT 14
#define FS_TIP 15
#define FS_THICK 16
#define FS_ROUGH 17
#define FS_SWEEP 18
#define FS_MIDSWEEP 19
#define FS_PLANFORM 20
#define FS_WETTED 21
#define FS_AR 22
#define FS_STALL 23
#define FS_BODY_R 24
#define FS_BODY_INTERF 25
#define FS_NUM_INTERF 26
#define FS_MAC_LEN 27
#define FS_MAC_Y 28
#define FS_XMAC_LE 29
#define FS_LE_SHAPE 30
#define FS_LE_THICK 31
#define FS_TE_THICK 32
#define FS_CP_POLY 33
#define FS_CANT 39
#define FS_NSLICES 40
#define FS_ACTUATED 41
#define FS_TIP_POS 42
#define FS_THICK_K 43
#define FS_SPANDIR 44
#define FS_MAX_FINS 8
#define FS_SLICE_R 60
#define FS_MAX_SLICES 16
#define FS_SLICE_A 76
#define FS_SLICE_LE 92
#define FS_SLICE_LEN 108
#define FS_CDTT_STAR 124
#define FS_CRIT_RE 125
#define FS_ROUGH_CF 126
/* per fin, constant while the fin is not actuated (Fins.py:474-482,532-551): deflected normal, axial (drag) direction,
   spanwise CP distance |span * (MACY + bodyRadius)| */
#define FS_FIN_NORMAL 128
#define FS_AXIAL_DIR 152
#define FS_SPANWISE_CP 176
/* Mach-independent pieces of the transonic blend (Fins.py:508-528 evaluates its four strip sums at the fixed Mach numbers 0.8, 0.801,
   1.4 and 1.401), computed once by the host with the reference's arithmetic */
#define FS_TRANS_CNA 184     /* CN_alpha (Fins.py:26-34) at Mach 0.8 and 0.8 + 0.001 */
#define FS_TRANS_K 186       /* Busemann K1, K2, K3 and Mach-cone slope (Fins.py:37-43,494-506) at Mach 1.4, then at 1.4 + 0.001 */
#define FS_TRANS_W 194       /* per-slice Mach-cone weight fractionOut + 0.5 fractionIn (CythonFinFunctions.pyx:92-110): 16 at 1.4, 16 at 1.401 */
/* flight-independent factors of the fin drag build-up (Fins.py:289-358) and of CN_alpha (Fins.py:26-34) */
#define FS_K_SKIN 226        /* FS_WETTED / Aref */
#define FS_K_THICKF 227      /* 1 + 2 FS_THICK / FS_MAC_LEN */
#define FS_K_LE 228          /* FS_LE_THICK FS_SPAN cos^2(FS_SWEEP) / Aref */
#define FS_TC 229            /* FS_THICK / FS_ROOT */
#define FS_K_PLAN 230        /* FS_PLANFORM / Aref */
#define FS_COS_MID 231       /* cos(FS_MIDSWEEP) and its reciprocal */
#define FS_RCOS_MID 232
#define FS_AR2 233           /* 2 FS_SPAN^2 / FS_PLANFORM */
#define FS_K_THICK_SUP 234   /* 2.3 FS_AR (FS_THICK / FS_MAC_LEN)^2, thickness drag above Mach 1 */
#define FS_SIZE 235

/* tabulated motor (Rocket/Propulsion.py:8-128). Table: 10 columns of MT_NROWS each, column-major,
   starting at C_P+MT_TABLE: time, thrust, oxWeight, fuelWeight, oxMOI xyz, fuelMOI xyz */
#define MT_NROWS 12
#define MT_OX_CG0 13
#define MT_OX_CG1 14
#define MT_FUEL_CG0 15
#define MT_FUEL_CG1 16
#define MT_OX_W0 17
#define MT_FUEL_W0 18
#define MT_TABLE 19
/* columns 10-13 serve per-lane motor adjust factors (LANE_MOTOR_ADJUST): time and thrust as read from the motor file, before
   impulseAdjustFactor / burnTimeAdjustFactor (Propulsion.py:109-111), and the oxidiser / fuel flow rates the propellant weights are
   integrated from (:113-128). Columns 0-3 hold the same quantities with the definition's own factors applied. */
#define MT_NCOLS 14
#define MT_COL_RAW_TIME 10
#define MT_COL_RAW_THRUST 11
#define MT_COL_OX_FLOW 12
#define MT_COL_FUEL_FLOW 13

/* recovery system (Rocket/Recovery.py:10-48): up to 4 chute stages of 5 doubles:
   trigger event type, trigger value, chute area, Cd, delay */
#define RC_NSTAGES 12
#define RC_STAGE0 13
#define RC_STAGE_STRIDE 5
#define RC_MAX_STAGES 4
#define RC_SIZE 33

/* force-only and table-driven components (Rocket/RocketComponents.py:183-413). Their `position` is NOT offset by the
   stage position (the reference reads it straight from the dictionary, :219, :270). */
/* FixedForce (:183-204): constant force at C_POS plus a constant moment */
#define FF_FORCE 12
#define FF_MOMENT 15
#define FF_SIZE 18
/* AeroForce (:206-231): constant Cd, Cl, CMx, CMy, CMz */
#define AF_AREF 12
#define AF_LREF 13
#define AF_COEFFS 14
#define AF_SIZE 19
/* AeroDamping (:233-261): damping-coefficient vectors for the x, y, z moment coefficients (position is the origin) */
#define AD_AREF 12
#define AD_LREF 13
#define AD_XCOEFFS 14
#define AD_YCOEFFS 17
#define AD_ZCOEFFS 20
#define AD_SIZE 23
/* TabulatedAeroForce (:263-345), one key column: keys[TA_NROWS], then TA_NVALS value columns of TA_NROWS each
   (column-major) starting at TA_TABLE; TA_COEFF_IDX[k] = position of value column k in (CD, CL, CMx, CMy, CMz) */
#define TA_AREF 12
#define TA_LREF 13
#define TA_KEY 14
#define TA_NROWS 15
#define TA_NVALS 16
#define TA_COEFF_IDX 17
#define TA_NKEYS 22       /* key columns of the table; 1: TA_KEY + the 1-D table at TA_TABLE (MAPLEAF.Motion.linInterp) */
#define TA_KEYS 23        /* >1: the AKEY_* id of every key column, in column order (up to ND_MAX_DIM) */
#define TA_ND_OFF 29      /* >1: model-block offset of the scattered-interpolation record (ND_*), behind H_SHARED_LEN in global memory */
#define TA_TABLE 30
#define AKEY_MACH 0
#define AKEY_ALTITUDE 1
#define AKEY_UNITRE 2
#define AKEY_TOTALAOA 3
#define AKEY_ROLLANGLE 4
#define AKEY_AOA 5
#define AKEY_AOSS 6
/* TabulatedInertia (:347-376): times[TI_NROWS], then mass, CGx, CGy, CGz, MOIx, MOIy, MOIz columns (column-major) */
#define TI_NROWS 12
#define TI_TABLE 13
/* FractionalJetDamping (:378-413) */
#define JD_FRACTION 12
#define JD_SIZE 13

/* ---- control system record at H_CTRL_OFF (0: none): GNC/ControlSystems.py:45-211, MomentControllers.py:19-83, Actuators.py:43-139,
   actuatedSystem.py:25-66 ------------------------------------------------------------------------------------------------------- */
#define CS_UPDATE_DT 0        /* controlTimeStep = 1 / updateRate (0: run every step) */
#define CS_TARGET_Q 1         /* Stabilizer target orientation (Navigation.py:15-24) */
#define CS_MC_TYPE 5          /* 0: constant gains, 1: gains scheduled by a table */
#define CS_GAINS 6            /* Pxy Ixy Dxy Pz Iz Dz */
#define CS_GAIN_NKEYS 12
#define CS_GAIN_KEYS 13       /* AKEY_* of each gain-table key column */
#define CS_GAIN_INTERP_OFF 17
#define CS_STAGE 18           /* controlled fin set: stage index, index in the stage's component list */
#define CS_COMP 19
#define CS_NACT 20
#define CS_ACT_TAU 21         /* first-order actuator response time */
#define CS_ACT_HASMIN 22
#define CS_ACT_MIN 23
#define CS_ACT_HASMAX 24
#define CS_ACT_MAX 25
#define CS_DEFL_NKEYS 26      /* aero keys in front of the three desired-moment keys of the deflection table */
#define CS_DEFL_KEYS 27
#define CS_DEFL_INTERP_OFF 31
#define CS_FORCE_RK4 32       /* adaptive method + fixed update rate: RK4 during the controlled ascent (ControlSystems.py:126-135) */
#define CS_TABLEAU2_OFF 33
#define CS_SIZE 36
#define CS_MAX_ACT 8
/* N-D scattered linear interpolation (scipy LinearNDInterpolator + NearestNDInterpolator, Motion/Interpolation.py:109-124):
   Delaunay triangulation computed on the host with scipy. Data after the 4 counts: points[npts][dim], values[npts][nval],
   simplices[nsimp][dim+1] (point indices), transform[nsimp][dim+1][dim] (barycentric transform of each simplex, NaN if degenerate),
   neighbors[nsimp][dim+1] (scipy's Delaunay.neighbors: the simplex opposite vertex j, -1 on the hull; used by the device's directed walk) */
#define ND_DIM 0
#define ND_NVAL 1
#define ND_NPTS 2
#define ND_NSIMP 3
#define ND_DATA 4
#define ND_MAX_DIM 6
#define ND_MAX_VAL 8

/* ---- per-lane input arrays (mlb_set_lanes arrayIds) ----------------------------------- */
#define LANE_WIND 0          /* 3: Environment.ConstantMeanWind.velocity (ENU)        */
#define LANE_INIT_STATE 1    /* 13: pos, vel, quat, body-frame angular velocity         */
#define LANE_PINK_SEED 2     /* 3: integer seeds of the pink-noise generators           */
#define LANE_MOTOR_ADJUST 3  /* 8: impulseAdjustFactor, burnTimeAdjustFactor of the motor of stage 0, 1, 2, 3 (definition order)   */
#define LANE_START 4         /* 2: start time and first step size (dropped stages start at their separation, SingleSimulations.py:280-299) */
#define LANE_NUM_ARRAYS 5
/* One row per stage separation of a lane (mlb_stage_separations): the state the dropped stage starts from */
#define SEP_STATE 0          /* 13: pos, vel, quat, body-frame angular velocity of the whole rocket at the separation                 */
#define SEP_TIME 13
#define SEP_DT 14            /* the step size the time loop was using (SimulationRunners' dts[0] at that moment)                        */
#define SEP_VALID 15         /* 1 when the separation happened                                                                          */
#define SEP_W 16
#define MLB_MAX_STAGES 4

/* ---- per-lane outputs -------------------------------------------------------------------- */
#define RES_LAND_X 0
#define RES_LAND_Y 1
#define RES_LAND_Z 2
#define RES_APOGEE 3
#define RES_MAX_SPEED 4
#define RES_FLIGHT_TIME 5
#define RES_MAX_HVEL 6
#define RES_SIZE 7
/* final-state rows: pos3 vel3 quat4 angvel3 time */
#define FIN_SIZE 14

/* lane status */
#define ST_OK 0
#define ST_RUNNING 1
#define ST_STEP_LIMIT 2
#define ST_NAN 3
#define ST_UNSUPPORTED 4

#endif
