// sampling_server.cpp — pybind11 entry of the sampling server: `sampling_server.Run(fanout, gpu_number,
// in_memory_mode, cache_mode)`, the in-process twin of `build/bin/sampling_server <gpu_number> <cache_agg_mode>`.
// Mirrors the reference module (sampling_server/sampling_server.cpp:7-22): same module name, same function, same
// argument order and meaning; reads ./meta_config from the cwd like the binary.  Unlike the reference, the fan-out
// passed here IS honoured (the reference forwards it to Server::Initialize too, engine/server.cu:46-88).
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <iostream>
#include <vector>

#include "server.h"

static int Run(const std::vector<int>& fanout, int gpu_number, int in_memory_mode, int cache_mode) {
  std::cout << "Start Sampling Server\n";
  Server* server = NewGPUServer();
  {
    pybind11::gil_scoped_release release;  // the serving loop blocks on semaphores; let other Python threads run
    server->Initialize(gpu_number, fanout, in_memory_mode);  // gpu number, default 1; in memory, default true
    server->PreSc(cache_mode);                               // cache aggregate mode, default 0
    server->Run();
    server->Finalize();
  }
  return 0;
}

PYBIND11_MODULE(sampling_server, m) {
  m.doc() = "Legion sampling server (B200 build): Run(fanout, gpu_number, in_memory_mode, cache_mode)";
  m.def("Run", &Run, "Run Sampling Server", pybind11::arg("fanout"), pybind11::arg("gpu_number") = 1,
        pybind11::arg("in_memory_mode") = 1, pybind11::arg("cache_mode") = 0);
}
