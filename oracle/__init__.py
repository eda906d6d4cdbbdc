"""CPU oracle for the Stratified-Transformer window-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`stratified_transformer_b200/`) imports this directory.  The only permitted
callers are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and only as the checker / timed CPU
baseline, never as the shipped path.

Contents
--------
attention_oracle.py  gather / index_add restatement of the attention math
                     (reference kernels: lib/pointops2/src/attention*/, rpe*/;
                     authors' torch form: lib/pointops2/functions/
                     test_relative_pos_encoding_op_step2.py:32-38)
index_oracle.py      voxel_grid / grid_sample / get_indice_pairs / CSR / rel-pos
                     index restatement (model/stratified_transformer.py:10-65,
                     186-188,267-317; model/swin3d_transformer.py:129-154)
fps_oracle.c/.py     exact furthest point sampling restatement
                     (lib/pointops2/src/sampling/sampling_cuda_kernel.cu:5-129)
ref_cuda.py          ctypes binding of oracle/_ref/libpointops2_ref.so, i.e. the
                     reference's OWN .cu files compiled where they lie (GPU only)
Makefile             builds oracle/_ref/ (needs /root/reference) and the C oracle

Pinning status: see the header of each module and DESIGN.md section "Oracle".
"""
