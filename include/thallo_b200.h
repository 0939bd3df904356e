/* thallo_b200 extension entry points (C ABI, plain pointers and sizes).
 *
 * These are the seam the north-star names: "per-energy residual and partial-derivative
 * device functions are emitted as CUDA C++ and JIT-compiled for sm_100a (NVRTC through a
 * thin C-ABI shim called from the host)".  In the reference the equivalent hand-over is
 * in-process Lua: ProblemSpec:Functions registers the generated `fmap` tables
 * (API/src/thallo.t:444-456,3468-3501) and gauss_newton.t:115 consumes them, then
 * util.makeGPUFunctions hands the kernels to terralib.cudacompile (util.t:797-927,
 * cuda_util.t:470).  A Terra/Lua (or Python) host calls ThalloB200_ProblemDefineFromSource
 * instead of that path; everything after it is the unchanged Thallo.h API.
 */
#pragma once
#include "Thallo.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Define a problem from an already-lowered energy: `descriptor` is the line-based plan
 * descriptor and `cuda_source` the generated translation unit (both NUL-terminated text,
 * see DESIGN.md "front-end seam").  Replaces thallo.t:444-456 + util.t:868. */
Thallo_Problem* ThalloB200_ProblemDefineFromSource(Thallo_State* state, const char* descriptor,
                                                   const char* cuda_source, const char* solverkind);

/* Compile only: NVRTC the source for sm_100a without touching a GPU.  Returns 0 on success,
 * and the size of the cubin in *cubin_size; the log (NUL-terminated) is copied into log
 * (up to log_capacity bytes).  Used by the CPU-side build check. */
int ThalloB200_CompileOnly(const char* cuda_source, char* log, unsigned long log_capacity, unsigned long* cubin_size);

/* The caller's stream (a cudaStream_t) the solver is ordered with; default is the legacy default stream like the
 * reference (util.t:769-772).  The solver's own work runs on a private stream per plan (so that the PCG iterations can
 * be captured into CUDA graphs): every entry point makes that stream wait for the caller's stream and every exit makes
 * the caller's stream wait for it, i.e. the caller sees plain stream-ordered behaviour on its stream. */
void ThalloB200_SetStream(Thallo_State* state, void* cuda_stream);

/* Number of kernel launches issued by this plan since creation (for bench accounting). */
unsigned long long ThalloB200_PlanLaunchCount(Thallo_State* state, Thallo_Plan* plan);

/* Linear (PCG) iterations executed in the most recent nonlinear step and in total. */
int ThalloB200_PlanLastLinearIterations(Thallo_State* state, Thallo_Plan* plan);
unsigned long long ThalloB200_PlanTotalLinearIterations(Thallo_State* state, Thallo_Plan* plan);

/* Copy solver vector `name` ("delta","r","p","Ap_X","preconditioner","z","b","CtC") to host memory
 * (count reals of the plan's precision).  Test hook. Returns number of reals copied. */
long long ThalloB200_PlanReadVector(Thallo_State* state, Thallo_Plan* plan, const char* name, void* host_dst, long long count);

/* Device pointer of solver vector `name` (nunk reals of the plan's precision; valid until Thallo_PlanFree), or NULL.
 * Measurement hook: lets a checker form dot products of full-size vectors on the device instead of copying them out. */
void* ThalloB200_PlanVectorPointer(Thallo_State* state, Thallo_Plan* plan, const char* name);

/* Materialise the Jacobian of residual group `group` at the current unknowns (after Thallo_ProblemInit), in the
 * reference's CSR order (precomputeJ / generateDumpJ, gauss_newton.t:325-487,1019-1025): element-major, within an
 * element row by row, row k holding the descriptor's row_nnz[k] entries; host_vals receives reals of the plan's
 * precision, host_cols the flat index of the unknown scalar each partial derivative belongs to (imageOffset +
 * channels*idx + channel) or -1 where the access falls outside the domain.  Returns the number of entries
 * (count * nnz per element) or -1.  Test / export hook: the solver itself never forms this copy. */
long long ThalloB200_PlanExportJacobian(Thallo_State* state, Thallo_Plan* plan, int group, void* host_vals,
                                        long long* host_cols, long long capacity);

/* Per-kernel device times accumulated since plan creation when the state was created with
 * timingLevel >= 2 (the reference wraps every launch in an event pair at that level,
 * util.t:774-790).  Writes "kernel_name launches total_ms\n" lines; returns bytes written. */
long long ThalloB200_PlanKernelTimes(Thallo_State* state, Thallo_Plan* plan, char* buf, long long capacity);

/* ---- multi-GPU: slab partition of an image / volume domain along its slowest axis.
 * One process per GPU.  Every rank lowers the energy for its LOCAL extent including the ghost
 * layers it shares with its neighbours (descriptor line "partition <ghost_lo> <ghost_hi>", see
 * thallo_b200/distributed.py) and passes local slabs of every image (ghost layers included).  The
 * reference has no multi-device path (SURVEY 2b "Collectives: none").
 *   ThalloB200_NcclUniqueId       rank 0 creates the NCCL id (128 bytes); the host program distributes it
 *   ThalloB200_PlanInitComm       collective: NCCL communicator for the PCG scalar reductions
 *   ThalloB200_PlanIpcHandle      CUDA IPC handle (64 bytes) of this plan's solver vectors + local slow-axis extent
 *   ThalloB200_PlanConnect        map the neighbours' solver vectors (NULL = no neighbour on that side); halo layers
 *                                 of p / z / delta are then stored straight into the neighbours' memory over NVLink
 * All return 0 on success. */
int ThalloB200_NcclUniqueId(void* id, int capacity);
int ThalloB200_PlanInitComm(Thallo_State* state, Thallo_Plan* plan, const void* nccl_id, int rank, int world);
int ThalloB200_PlanIpcHandle(Thallo_State* state, Thallo_Plan* plan, void* handle64, long long* slow_extent);
int ThalloB200_PlanConnect(Thallo_State* state, Thallo_Plan* plan, const void* handle_lo, long long extent_lo,
                           const void* handle_hi, long long extent_hi);

/* ---- multi-GPU: vertex partition of a graph domain (energies over a vertex domain and an edge
 * domain, gather schedule).  Every rank owns a contiguous vertex range and lowers the energy for its LOCAL
 * vertices (ghost copies of the neighbouring ranks' vertices in front of and behind the owned range) and its
 * local edges (the edges it owns first, then the foreign edges that touch its vertices); descriptor line
 * "gpartition", see thallo_b200/distributed.py graph_partition.  Communicator and IPC handle as above
 * (*slow_extent = number of local vertices); instead of ThalloB200_PlanConnect:
 *   ThalloB200_PlanConnectGraph   map the neighbours' solver vectors; width_lo / width_hi = how many of THIS rank's
 *                                 first / last owned vertices the lower / upper neighbour holds as ghosts.  The
 *                                 ghost values of p / z / delta are stored straight into the neighbours' memory. */
int ThalloB200_PlanConnectGraph(Thallo_State* state, Thallo_Plan* plan, const void* handle_lo, long long extent_lo,
                                long long width_lo, const void* handle_hi, long long extent_hi, long long width_hi);

/* ---- multi-GPU: all-to-all peer mapping (replaces ThalloB200_PlanConnect / ThalloB200_PlanConnectGraph).
 * Every rank maps EVERY other rank's solver-vector block.  After this call the PCG iteration of a partitioned plan
 * issues no collective and no helper kernel: the kernel that produces a dot product all-reduces it itself -- its last
 * CTA stores the rank's partial sums into a mailbox in every peer's memory over NVLink, waits for the peers' values
 * in its own mailbox and adds them up in rank order (bit-identical on all ranks) -- and the kernel that computes z
 * stores its boundary layers straight into the neighbours' ghost layers, ordered before its mailbox flag.  NCCL
 * remains for the once-per-nonlinear-iteration scalars (cost, model cost) and for vector all-reduces (the replicated
 * camera block of bundle adjustment).  THALLO_B200_MG_NCCL=1 keeps the NCCL sequence for comparison.
 *   ThalloB200_PlanPeerInfo     info4 = {local extent of the partitioned axis, bytes between consecutive solver
 *                               vectors, ghost_lo, ghost_hi} of this rank's plan
 *   ThalloB200_PlanConnectAll   handles64 = world x 64-byte IPC handles (ThalloB200_PlanIpcHandle) in rank order,
 *                               infos4 = world x 4 values of ThalloB200_PlanPeerInfo; call after ThalloB200_PlanInitComm */
int ThalloB200_PlanPeerInfo(Thallo_State* state, Thallo_Plan* plan, long long* info4);
int ThalloB200_PlanConnectAll(Thallo_State* state, Thallo_Plan* plan, int world, const void* handles64, const long long* infos4);

/* Known-answer tests of the warp primitives behind the residualwise scatter path (ballot, peer
 * discovery by key, by-key reduction before the atomic), written after the reference's
 * tests/cuda_unit_tests/{ballot,get_peers,reduce_peers}.t; one warp each on the current device.
 *   which = 0  ballot:       out[0] = max over lanes of ballot(lane id)            (reference asserts 0xfffffffe)
 *   which = 1  get_peers:    out[0] = sum over lanes of (peers(lane % 4) & 0xff)   (reference asserts 255*32/4)
 *   which = 2  reduce_peers: out[k] = float sums, out[nkeys + k] = double sums of the lane ids with lane % nkeys == k
 *                                                                                  (reference, nkeys = 4: 112 + 8 k)
 * `out` receives up to `capacity` doubles.  Returns the number written, or -1 on failure. */
int ThalloB200_WarpSelfTest(int which, int nkeys, double* out, int capacity);

/* Last error message of this thread ("" if none). */
const char* ThalloB200_LastError(void);

/* Library version string. */
const char* ThalloB200_Version(void);

#ifdef __cplusplus
}
#endif
