// Conditional section of a captured launch sequence (CUDA graph IF node).
// The attack loop replays ONE CUDA graph per iteration; in an iteration in which no image takes the network branch
// (attack_rd.py:333-334 -- with the reference's schedule that is ~99 % of the iterations) every network launch used to be
// replayed and exit at once: ~0.2 ms of the 0.75 ms such an iteration costs at 64 images, and 21 of the 26 launches of the
// launch-latency-bound small-batch case (train.py --adv, 8 x 256x256: 95 us per iteration).  With this pair of calls the
// caller puts those launches into the body of an IF node whose condition is read from device memory at replay time.
#include <cuda_runtime.h>

#include "icadv_common.cuh"

namespace icadv {

__global__ void graph_if_set_kernel(cudaGraphConditionalHandle handle, const int* __restrict__ flag) {
  cudaGraphSetConditional(handle, *flag > 0 ? 1u : 0u);
}

}  // namespace icadv

using namespace icadv;

extern "C" {

int icadv_graph_if_begin(const int* flag, icadv_stream_t capture_stream, icadv_stream_t body_stream) {
  ICADV_REQUIRE(flag && capture_stream && body_stream && capture_stream != body_stream, "bad graph_if_begin args");
  cudaStream_t cap = as_stream(capture_stream), body = as_stream(body_stream);
  cudaStreamCaptureStatus status;
  unsigned long long id;
  cudaGraph_t graph;
  const cudaGraphNode_t* deps;
  size_t n_deps;
  ICADV_CUDA_TRY(cudaStreamGetCaptureInfo(cap, &status, &id, &graph, &deps, &n_deps));
  ICADV_REQUIRE(status == cudaStreamCaptureStatusActive, "graph_if_begin: the stream is not capturing");
  cudaGraphConditionalHandle handle;
  ICADV_CUDA_TRY(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
  graph_if_set_kernel<<<1, 1, 0, cap>>>(handle, flag);
  ICADV_CUDA_TRY(cudaGetLastError());
  ICADV_CUDA_TRY(cudaStreamGetCaptureInfo(cap, &status, &id, &graph, &deps, &n_deps));
  cudaGraphNodeParams params = {};
  params.type = cudaGraphNodeTypeConditional;
  params.conditional.handle = handle;
  params.conditional.type = cudaGraphCondTypeIf;
  params.conditional.size = 1;
  cudaGraphNode_t node;
  ICADV_CUDA_TRY(cudaGraphAddNode(&node, graph, deps, n_deps, &params));
  // everything captured on `capture_stream` from here on depends on the IF node
  ICADV_CUDA_TRY(cudaStreamUpdateCaptureDependencies(cap, &node, 1, cudaStreamSetCaptureDependencies));
  ICADV_CUDA_TRY(cudaStreamBeginCaptureToGraph(body, params.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                               cudaStreamCaptureModeRelaxed));
  return ICADV_OK;
}

int icadv_graph_if_create(icadv_stream_t capture_stream, unsigned long long* handle_out) {
  ICADV_REQUIRE(capture_stream && handle_out, "bad graph_if_create args");
  cudaStreamCaptureStatus status;
  unsigned long long id;
  cudaGraph_t graph;
  const cudaGraphNode_t* deps;
  size_t n_deps;
  ICADV_CUDA_TRY(cudaStreamGetCaptureInfo(as_stream(capture_stream), &status, &id, &graph, &deps, &n_deps));
  ICADV_REQUIRE(status == cudaStreamCaptureStatusActive, "graph_if_create: the stream is not capturing");
  cudaGraphConditionalHandle handle;
  ICADV_CUDA_TRY(cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault));
  *handle_out = (unsigned long long)handle;
  return ICADV_OK;
}

int icadv_graph_if_begin_handle(unsigned long long handle, icadv_stream_t capture_stream, icadv_stream_t body_stream) {
  ICADV_REQUIRE(handle && capture_stream && body_stream && capture_stream != body_stream, "bad graph_if_begin_handle args");
  cudaStream_t cap = as_stream(capture_stream), body = as_stream(body_stream);
  cudaStreamCaptureStatus status;
  unsigned long long id;
  cudaGraph_t graph;
  const cudaGraphNode_t* deps;
  size_t n_deps;
  ICADV_CUDA_TRY(cudaStreamGetCaptureInfo(cap, &status, &id, &graph, &deps, &n_deps));
  ICADV_REQUIRE(status == cudaStreamCaptureStatusActive, "graph_if_begin_handle: the stream is not capturing");
  cudaGraphNodeParams params = {};
  params.type = cudaGraphNodeTypeConditional;
  params.conditional.handle = (cudaGraphConditionalHandle)handle;
  params.conditional.type = cudaGraphCondTypeIf;
  params.conditional.size = 1;
  cudaGraphNode_t node;
  ICADV_CUDA_TRY(cudaGraphAddNode(&node, graph, deps, n_deps, &params));
  ICADV_CUDA_TRY(cudaStreamUpdateCaptureDependencies(cap, &node, 1, cudaStreamSetCaptureDependencies));
  ICADV_CUDA_TRY(cudaStreamBeginCaptureToGraph(body, params.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                               cudaStreamCaptureModeRelaxed));
  return ICADV_OK;
}

int icadv_graph_if_end(icadv_stream_t body_stream) {
  ICADV_REQUIRE(body_stream != nullptr, "bad graph_if_end args");
  cudaGraph_t g = nullptr;
  ICADV_CUDA_TRY(cudaStreamEndCapture(as_stream(body_stream), &g));
  return ICADV_OK;
}

}  // extern "C"
