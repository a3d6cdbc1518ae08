/*
 * unitair_b200.h -- C ABI of the B200-native state-vector engine that sits behind
 * unitair's gate-application API.
 *
 * The reference (qcware/qcware-unitair v0.3.0) is pure Python over PyTorch and has no
 * FFI of its own; the boundary it exposes for this path is the Python function surface
 * of src/unitair/simulation/operations.py and src/unitair/states/innerprod.py.  Each
 * entry point below replaces the ATen work behind one of those functions (cited per
 * function).  The host shim (qcware-unitair_b200/unitair_b200, Python like the
 * reference) keeps the reference's signatures, validation and error types and calls
 * these through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless named host_*.  No torch types cross the ABI.
 *  - A state is `batch` contiguous rows of 2^num_qubits complex numbers (interleaved
 *    re,im).  Qubit q is bit (num_qubits-1-q) of the row index: qubit 0 is the MOST
 *    significant bit (src/unitair/states/conversions.py:43-45).
 *  - A gate is row-major (2^k x 2^k); gate index MSB <-> qubits[0]
 *    (src/unitair/simulation/operations.py:82-86).
 *  - dtype: UA_C64 = complex64 (2 x f32), UA_C128 = complex128 (2 x f64).
 *  - Every function returns UA_OK (0) or an error code, never throws; ua_last_error()
 *    gives the message.  Work is enqueued on `stream` (a cudaStream_t) and is
 *    asynchronous; outputs/workspaces are allocated by the caller.
 *  - State pointers must be 16-byte aligned.
 */
#ifndef UNITAIR_B200_H
#define UNITAIR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UA_OK 0
#define UA_ERR_INVALID 1      /* bad argument (caller bug)            */
#define UA_ERR_UNSUPPORTED 2  /* valid request this build cannot run  */
#define UA_ERR_CUDA 3         /* CUDA runtime error at launch         */

#define UA_C64 0
#define UA_C128 1

#define UA_MAX_GATE_QUBITS 5        /* register/shared-memory kernels            */
#define UA_MAX_GENERIC_GATE_QUBITS 10 /* slow generic kernel (out-of-place only) */
#define UA_MAX_FUSED_GATES 64       /* gates per fused shared-memory pass        */
#define UA_MAX_TILE_BITS 14

/* library info ----------------------------------------------------------------- */
int ua_version(void);
const char *ua_last_error(void);
/* number of kernels this library has launched so far in this process */
unsigned long long ua_launch_count(void);

/* Dense k-qubit gate, out = U . in (or U^H . in when adjoint != 0).
 * Replaces apply_operator_tensor / act_first_qubits_tensor / permute_qubits_tensor
 * (src/unitair/simulation/operations.py:151-186, 258-329, 626-654): no permute copies,
 * target strides are computed in-kernel.
 *   out            batch x 2^n, may alias `in` (in-place) when in_batch_stride == 2^n
 *   in             state(s); row b starts at in + b*in_batch_stride (0 = one state
 *                  broadcast to every gate of the batch, operations.py:304-309)
 *   gate           row-major 2^k x 2^k; matrix b at gate + b*gate_batch_stride
 *                  (0 = one gate shared by the whole batch)
 *   host_qubits    k distinct unitair qubit indices in [0, n), in gate order       */
int ua_apply_gate(int dtype, void *out, const void *in, const void *gate,
                  int num_qubits, int k, const int *host_qubits,
                  long long batch, long long in_batch_stride, long long gate_batch_stride,
                  int adjoint, void *stream);

/* Gradient of a real loss w.r.t. the gate (PyTorch's conjugate-Wirtinger convention;
 * autograd of operations.py:322):
 *   grad_gate[a,b] = sum_r grad_out[a,r] * conj(psi_in[b,r])
 * per batch entry when gate_batch_stride != 0, summed over the batch when it is 0.
 * workspace: ua_gate_grad_workspace_bytes() bytes of scratch.                        */
size_t ua_gate_grad_workspace_bytes(int dtype, int num_qubits, int k, long long batch,
                                    long long gate_batch_stride);
int ua_gate_grad(int dtype, void *grad_gate, const void *grad_out, const void *psi_in,
                 int num_qubits, int k, const int *host_qubits,
                 long long batch, long long psi_batch_stride, long long gate_batch_stride,
                 void *workspace, size_t workspace_bytes, void *stream);

/* Fused diagonal phase: out[b,e] = exp(-+ i angles[b*abs + e*aes]) * in[b*ibs + e].
 * Replaces torch.exp(-1j*angles)*state (operations.py:41-42) in one pass.
 * conj_phase != 0 gives exp(+i angle) (the backward of apply_phase).
 * angles are f32 for UA_C64 and f64 for UA_C128.                                     */
int ua_apply_phase(int dtype, void *out, const void *in, const void *angles,
                   long long elems, long long batch, long long in_batch_stride,
                   long long angle_batch_stride, long long angle_elem_stride,
                   int conj_phase, void *stream);

/* Backward of apply_phase in one pass:
 *   grad_state[b,e]  = exp(+i angle) * grad_out[b,e]
 *   grad_angle[b,e]  = Im( conj(grad_out[b,e]) * exp(-i angle) * psi_in[b,e] )  (real)
 * grad_state / grad_angle may be NULL to skip.                                       */
int ua_phase_backward(int dtype, void *grad_state, void *grad_angle, const void *grad_out,
                      const void *psi_in, const void *angles,
                      long long elems, long long batch, long long in_batch_stride,
                      long long angle_batch_stride, long long angle_elem_stride,
                      void *stream);

/* Reductions (src/unitair/states/innerprod.py:4-65), one read of the state each.
 * Results are written in the state's real/complex precision.
 * workspace: ua_reduce_workspace_bytes(batch, elems) bytes of scratch.               */
size_t ua_reduce_workspace_bytes(long long batch, long long elems);
int ua_abs_squared(int dtype, void *out_real, const void *in, long long count, void *stream);
int ua_norm_squared(int dtype, void *out_real, const void *in, long long elems, long long batch,
                    void *workspace, size_t workspace_bytes, void *stream);
int ua_diag_expectation(int dtype, void *out_real, const void *diag_real, const void *in,
                        long long elems, long long batch, long long diag_batch_stride,
                        long long in_batch_stride,
                        void *workspace, size_t workspace_bytes, void *stream);
int ua_inner_product(int dtype, void *out_complex, const void *a, const void *b,
                     long long elems, long long batch, long long a_batch_stride,
                     long long b_batch_stride,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Backward of abs_squared / norm_squared / diag_expectation_value w.r.t. the state in one pass
 * (autograd of src/unitair/states/innerprod.py:26,46,59; the eager formula 2*g*d*psi costs three
 * kernels and two full-size temporaries):
 *   out[b,e] = scale * g[b*g_batch_stride + e*g_elem_stride] * (diag ? diag[b*diag_batch_stride + e] : 1)
 *              * in[b*in_batch_stride + e]
 * g / diag are real (f32 for UA_C64, f64 for UA_C128); strides are 0 or the natural ones.        */
int ua_real_scale(int dtype, void *out, const void *in, const void *g_real, const void *diag_real,
                  long long elems, long long batch, long long in_batch_stride,
                  long long g_batch_stride, long long g_elem_stride, long long diag_batch_stride,
                  double scale, void *stream);

/* Fused shared-memory pass: a list of dense gates (k <= 3 each) whose target bits all
 * lie inside one tile = the low `tile_low_bits` index bits plus `num_high` chosen higher
 * bit positions.  Every tile is staged once in shared memory, all gates are applied
 * there, and the tile is written back: one HBM read+write for the whole list.
 * This is the engine behind apply_all_qubits (operations.py:332-413) and the circuit
 * API (apply_to_qubits-style fusion, operations.py:416-503).
 *   total_bits        log2 of the flat index space covered (num_qubits [+ batch bits]);
 *                     rows = total_amps >> total_bits independent spaces
 *   host_high_pos     ascending bit positions (>= tile_low_bits) that join the tile
 *   host_gate_k[g]    qubits of gate g; host_gate_bits[g*3+j] its BIT POSITIONS
 *                     (gate order, MSB first); gate_mats: packed matrices, matrix g at
 *                     gate_mats + host_gate_offset[g] (+ row*gate_row_stride if != 0) */
int ua_apply_fused_pass(int dtype, void *out, const void *in, long long total_amps,
                        int total_bits, int tile_low_bits, int num_high,
                        const int *host_high_pos, int num_gates, const int *host_gate_k,
                        const int *host_gate_bits, const long long *host_gate_offset,
                        const void *gate_mats, long long gate_row_stride, int adjoint,
                        void *stream);
/* Fused pass whose output is SCATTERED over 2^m destination buffers: the global-qubit exchange
 * of a state sharded over several GPUs folded into the last pass before it (the reference is
 * single-device; unitair_b200/sharded.py).  The m index bits host_scatter_pos[]
 * (ascending, >= tile_low_bits, none of them a tile bit) are removed from the output index:
 *   out_index = in_index with the scatter bits squeezed out,   buffer = value of those bits
 * (bit j of the buffer number = index bit host_scatter_pos[j]).  host_dst_ptrs[b] is the start
 * of destination block b (2^(total_bits-m) amplitudes) -- normally memory of a PEER GPU mapped
 * into this process (ua_ipc_open), so the tiles travel over NVLink while the pass computes.
 * num_gates may be 0 (pure scatter copy).  One state only (total_amps == 2^total_bits), shared
 * gates only.  The caller orders the pass against the peers' reads (a collective barrier).
 * visit_xor: tiles whose scatter bits have the value v are visited when the kernel's tile
 * counter reaches v ^ visit_xor.  With visit_xor = this rank's own block number every rank
 * writes to a different peer at any moment (no incast on one GPU's NVLink ingress).           */
#define UA_MAX_SCATTER_BITS 3
int ua_apply_fused_pass_scatter(int dtype, const void *in, long long total_amps, int total_bits,
                                int tile_low_bits, int num_high, const int *host_high_pos,
                                int num_gates, const int *host_gate_k, const int *host_gate_bits,
                                const long long *host_gate_offset, const void *gate_mats,
                                int num_scatter_bits, const int *host_scatter_pos,
                                void *const *host_dst_ptrs, int visit_xor, void *stream);

/* The same two passes for complex64 circuits of SHARED 1-/2-qubit gates whose matrix VALUES are
 * known on the host (host_gate_mats: host memory, interleaved re,im floats, same packing and
 * offsets as gate_mats above).  The matrices travel in the kernel parameters and stay in uniform
 * registers; the gates are applied to register-resident groups of 16 amplitudes ("clusters" of
 * four tile bits, one shared-memory round trip per cluster instead of per gate); three
 * 256-thread teams per SM work on a ring of tile buffers that a producer warp keeps in flight
 * to and from HBM (csrc/ua_cluster.cu).  Returns UA_ERR_UNSUPPORTED when the pass does not fit
 * this path (a gate with k > 2, complex128, a tile below 4 bits or above 2^12 amplitudes, more
 * than 5 TMA dimensions with the first one fixed to the 4 lowest bits): call the device-matrix
 * entry point instead.                                                                        */
int ua_apply_fused_pass_hostmats(int dtype, void *out, const void *in, long long total_amps,
                                 int total_bits, int tile_low_bits, int num_high,
                                 const int *host_high_pos, int num_gates, const int *host_gate_k,
                                 const int *host_gate_bits, const long long *host_gate_offset,
                                 const void *host_gate_mats, int adjoint, void *stream);
int ua_apply_fused_pass_scatter_hostmats(int dtype, const void *in, long long total_amps, int total_bits,
                                         int tile_low_bits, int num_high, const int *host_high_pos,
                                         int num_gates, const int *host_gate_k, const int *host_gate_bits,
                                         const long long *host_gate_offset, const void *host_gate_mats,
                                         int num_scatter_bits, const int *host_scatter_pos,
                                         void *const *host_dst_ptrs, int visit_xor, void *stream);

/* Peer memory for the scatter pass: export a device allocation of this process / map one of
 * another process on the same node (CUDA IPC).  ua_ipc_export writes a 64-byte handle for the
 * allocation containing `ptr` and the offset of `ptr` inside it; ua_ipc_open maps the peer's
 * allocation and returns the address corresponding to the peer's `ptr`; ua_ipc_close unmaps
 * (pass the pointer ua_ipc_open returned and the same offset).                                */
int ua_ipc_export(const void *ptr, unsigned char handle_out[64], long long *offset_out);
int ua_ipc_open(const unsigned char handle[64], long long offset, void **ptr_out);
int ua_ipc_close(void *ptr, long long offset);

/* limits of the fused pass for this dtype: largest tile (bits) and the total number of
 * complex matrix elements (sum of 4^k over the gates) one pass can hold */
int ua_fused_limits(int dtype, int *max_tile_bits_out, int *max_matrix_elems_out);

/* Fused BACKWARD pass of the adjoint method for a pass of 1-/2-qubit UNITARY gates (same tile
 * description as ua_apply_fused_pass, gates in FORWARD order).  In place on both buffers:
 *   psi  : state AFTER the pass on entry, state BEFORE the pass on return (U^H applied in reverse)
 *   grad : gradient w.r.t. the pass output on entry, w.r.t. its input on return
 * For every gate g with host_gate_needs_grad[g] != 0, sum_r grad_out[a,r] conj(psi_in[b,r]) is
 * ADDED to grad_acc (fp64 complex, zero-initialised by the caller, [rows or 1][sum 4^k] laid
 * out like the forward pass's shared-memory matrices: gate g at element offset sum_{g'<g} 4^k',
 * entry (s,t) in TARGET-BIT order: bit i of s <-> i-th lowest target bit position).            */
int ua_fused_backward_pass(int dtype, void *psi, void *grad, long long total_amps,
                           int total_bits, int tile_low_bits, int num_high,
                           const int *host_high_pos, int num_gates, const int *host_gate_k,
                           const int *host_gate_bits, const long long *host_gate_offset,
                           const void *gate_mats, long long gate_row_stride,
                           const int *host_gate_needs_grad, void *grad_acc, void *stream);

/* Bit-permutation of the index (qubit swap / permute, operations.py:506-654), one pass:
 * out[b, i] = in[b, j] where bit host_src_bit[p] of j = bit p of i.                   */
int ua_permute_bits(int dtype, void *out, const void *in, int num_bits, long long batch,
                    const int *host_src_bit, void *stream);

/* Sign-mask diagonal gates (CZ, CC...CZ and products), one pass, masks tested in-kernel:
 * out[b, i] = in[b, i] * prod_m (-1)^[(i & mask_m) == mask_m]; mask bit p <-> index bit p.
 * Replaces multi_cz / multi_controlled_z (operations.py:657-783).  out may alias in.         */
int ua_apply_sign_masks(int dtype, void *out, const void *in, int num_qubits, long long batch,
                        int num_masks, const unsigned long long *host_masks, void *stream);

/* Sampling in the computational basis (measure, src/unitair/simulation/measurement.py:9-70)
 * without materialising abs_squared(state) or torch.distributions.Categorical's temporaries:
 *   ua_sample_block_sums  out_f64[b] = sum of |in[i]|^2 over block b of 2^block_log2 amplitudes
 *                         (one read of the state, fp64 accumulation);
 *   ua_sample_locate      for every draw targets_f64[s] in [0, total) find the amplitude index i
 *                         with cdf(i-1) <= draw < cdf(i), given the INCLUSIVE cumulative sum of
 *                         the block sums (block_cdf_f64); out_index_i64[s] = i.
 * 5 <= block_log2 <= 20.  The cumulative sum of the (few) block sums and the uniform draws are
 * the caller's (tiny tensors).                                                                */
int ua_sample_block_sums(int dtype, void *out_f64, const void *in, long long elems,
                         int block_log2, void *stream);
int ua_sample_locate(int dtype, void *out_index_i64, const void *in, long long elems,
                     int block_log2, const void *block_cdf_f64, const void *targets_f64,
                     long long num_samples, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UNITAIR_B200_H */
