// Streaming engine (any B): placeholder entry points until the host-driven solver lands.
#include "phx_common.cuh"

extern "C" {
size_t phx_stream_workspace_bytes(const phx_ctx*, int, int, int, int, int) {
    phx_set_error("streaming engine not built yet");
    return 0;
}
int phx_stream_solve_forward(phx_ctx*, int, int, int, const float*, const float*, const double*, int, int, int, int,
                             double, double, int64_t, float*, void*, size_t, phx_status*, double*, int, void*) {
    phx_set_error("streaming engine not built yet");
    return PHX_ERR_UNSUPPORTED;
}
int phx_stream_solve_adjoint(phx_ctx*, int, int, int, const float*, const double*, int, int, int, double, double,
                             int64_t, const float*, const float*, float*, float*, void*, size_t, phx_status*, double*,
                             int, void*) {
    phx_set_error("streaming engine not built yet");
    return PHX_ERR_UNSUPPORTED;
}
}
