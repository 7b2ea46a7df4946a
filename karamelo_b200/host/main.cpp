// main.cpp - `kml -i script.mpm`: command-line entry, same flags as the reference
// (reference src/mpm.cpp:60-70: -i / -in <file>).
#include "sim.h"
#include <cstring>
#include <iostream>

int main(int argc, char **argv) {
  std::string file; int device = 0; bool quiet = false;
  for (int i = 1; i < argc; i++) {
    if ((!strcmp(argv[i], "-i") || !strcmp(argv[i], "-in")) && i + 1 < argc) file = argv[++i];
    else if (!strcmp(argv[i], "-d") && i + 1 < argc) device = atoi(argv[++i]);
    else if (!strcmp(argv[i], "-q")) quiet = true;
  }
  if (file.empty()) { std::cerr << "usage: kml -i <script.mpm> [-d device] [-q]\n"; return 2; }
  try {
    kmlh::Sim sim; sim.device = device; sim.quiet = quiet; sim.input.echo = !quiet;
    sim.logfile.open("log.mpm");
    std::cout << "backend: " << kml_backend() << std::endl;
    sim.input.file(file);
  } catch (const std::exception &e) { std::cerr << "ERROR: " << e.what() << std::endl; return 1; }
  return 0;
}
