// crixus_b200 -- command-line front end with the reference's process surface: `crixus <file.ini>`
// (/root/reference/src/crixus.cpp:7-10, src/crixus.cu:57-65).  Everything happens in cx_run_ini (frontend.cpp).
#include <cstdio>
#include "../../../include/crixus_b200.h"

int main(int argc, char **argv)
{
  if (argc > 2) printf("Ignoring additional arguments after configuration file.\n");
  // the reference's main() discards crixus_main's return code (crixus.cpp:7-10); we return it
  return cx_run_ini(argc > 1 ? argv[1] : nullptr) == CX_OK ? 0 : 1;
}
