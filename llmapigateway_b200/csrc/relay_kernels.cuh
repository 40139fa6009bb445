// placeholder until the bulk kernel lands: fast mode == exact sequential path
static inline cudaError_t launch_step_fast(const StepArgs& a, int sm_count, cudaStream_t stream, cudaEvent_t* ev, int* launched) {
    cudaError_t r;
    if ((r = cudaEventRecord(ev[1], stream)) != cudaSuccess) return r;
    if (a.n_bytes) { k_copy<<<sm_count * 8, 256, 0, stream>>>(a.data, a.out, a.n_bytes); ++*launched; }
    if ((r = cudaEventRecord(ev[2], stream)) != cudaSuccess) return r;
    if (a.n_segs) { k_general<<<(a.n_segs + 63) / 64, 64, 0, stream>>>(a); ++*launched; }
    if ((r = cudaEventRecord(ev[3], stream)) != cudaSuccess) return r;
    return cudaGetLastError();
}
