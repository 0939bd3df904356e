/* Thallo C ABI as exported by thallo_b200 (libThallo.so / libThallo.a).
 *
 * Drop-in replacement for reference API/release/include/Thallo.h:1-106: the same
 * twelve entry points, the same struct layouts, the same argument meaning.  A program
 * written against the reference header compiles and links against this one unchanged
 * (every tests/<name>/main.cpp, examples/shared/ThalloSolver.h:40-112).  Written from the
 * reference's documented interface; behaviour notes cite the reference implementation.
 */
#pragma once

#ifdef __cplusplus
extern "C" {
#endif

typedef struct Thallo_State Thallo_State;
typedef struct Thallo_Plan Thallo_Plan;
typedef struct Thallo_Problem Thallo_Problem;

/* Per-state options (reference Thallo.h:10-36, plumbing createwrapper.t:143-167).
 * A zeroed struct is the fast default. */
struct Thallo_InitializationParameters {
    int doublePrecision;  /* nonzero: unknowns and all solver vectors are double */
    int verbosityLevel;   /* 0 quiet, >=1 solver log */
    int timingLevel;      /* 0 none, 1 coarse buckets, >=2 also per-kernel events */
    int threadsPerBlock;  /* positive multiple of 32, else 256 (createwrapper.t:155-158) */
    int useAutoscheduler; /* accepted; thallo_b200 always picks the schedule itself */
    int cpuOnly;          /* must be 0: there is no CPU fallback in thallo_b200 */
};
typedef struct Thallo_InitializationParameters Thallo_InitializationParameters;

/* replaces Thallo.h:41 / createwrapper.t:130-221 */
Thallo_State* Thallo_NewState(Thallo_InitializationParameters params);

/* replaces Thallo.h:46-47 / thallo.t:93-99,5954-5961. solverkind is exactly
 * "gauss_newton" or "levenberg_marquardt" (thallo.t:74). */
Thallo_Problem* Thallo_ProblemDefine(Thallo_State* state, const char* filename, const char* solverkind);
void Thallo_ProblemDelete(Thallo_State* state, Thallo_Problem* problem);

/* replaces Thallo.h:52-53 / thallo.t:1384-1434,5963-5971. Returns NULL when the energy
 * cannot be compiled.  `dimensions` is read at plan time only. */
Thallo_Plan* Thallo_ProblemPlan(Thallo_State* state, Thallo_Problem* problem, unsigned int* dimensions);
void Thallo_PlanFree(Thallo_State* state, Thallo_Plan* plan);

/* replaces Thallo.h:57,61 / gauss_newton.t:1828-1862: value is read/written with the
 * parameter's own C type (float for the ten real-valued ones, int for
 * residual_reset_period, nIter, nIterations, lIterations); unknown names only warn. */
void Thallo_SetSolverParameter(Thallo_State* state, Thallo_Plan* plan, const char* name, void* value);
void Thallo_GetSolverParameter(Thallo_State* state, Thallo_Plan* plan, const char* name, void* value);

/* replaces Thallo.h:66,73,76 / thallo.t:5974-5986, gauss_newton.t:1166-1198,1545-1785.
 * problemparams[i]: device pointer for Unknown/Array/Sparse input i, host pointer to a
 * scalar for Param input i. */
void Thallo_ProblemSolve(Thallo_State* state, Thallo_Plan* plan, void** problemparams);
void Thallo_ProblemInit(Thallo_State* state, Thallo_Plan* plan, void** problemparams);
int Thallo_ProblemStep(Thallo_State* state, Thallo_Plan* plan, void** problemparams);

/* replaces Thallo.h:81 / gauss_newton.t:1787-1793 */
double Thallo_ProblemCurrentCost(Thallo_State* state, Thallo_Plan* plan);

struct Thallo_PerformanceEntry {
    unsigned int count;
    double minMS;
    double maxMS;
    double meanMS;
    double stddevMS;
};
typedef struct Thallo_PerformanceEntry Thallo_PerformanceEntry;

struct Thallo_PerformanceSummary {
    Thallo_PerformanceEntry total;
    Thallo_PerformanceEntry nonlinearIteration;
    Thallo_PerformanceEntry nonlinearSetup;
    Thallo_PerformanceEntry linearSolve;
    Thallo_PerformanceEntry nonlinearResolve;
};
typedef struct Thallo_PerformanceSummary Thallo_PerformanceSummary;

/* replaces Thallo.h:106 / gauss_newton.t:1795-1799, util.t:516-541 */
void Thallo_GetPerformanceSummary(Thallo_State* state, Thallo_Plan* plan, Thallo_PerformanceSummary* summary);

#ifdef __cplusplus
}
#endif
