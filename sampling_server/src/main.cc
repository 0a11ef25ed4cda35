// main.cc — `sampling_server <gpu_number> <cache_agg_mode>` (reference: sampling_server/src/main.cu:5-16).
// legion_server.py passes math.log2(clique) as e.g. "3.0"; atoi() keeps the integer part like the
// reference.  Fan-out: default {25,10} as hard-coded in the reference; LEGION_FANOUT="15,10,5" or
// extra meta_config fields override it.
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>

#include "server.h"

int main(int argc, char** argv) {
  std::cout << "Start Sampling Server\n";
  if (argc < 3) {
    std::cout << "usage: sampling_server <gpu_number> <cache_agg_mode>\n";
    return 1;
  }
  std::vector<int> fanout = {25, 10};
  if (const char* e = std::getenv("LEGION_FANOUT")) {
    fanout.clear();
    std::stringstream ss(e);
    std::string tok;
    while (std::getline(ss, tok, ',')) fanout.push_back(std::atoi(tok.c_str()));
  }
  Server* server = NewGPUServer();
  server->Initialize(std::atoi(argv[1]), fanout, 1);
  server->PreSc(std::atoi(argv[2]));
  server->Run();
  server->Finalize();
  return 0;
}
