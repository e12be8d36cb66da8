// main_meth — drop-in for the reference's main_meth.exe (src/main_meth.cpp): same command line, same output files.
#include "../../../include/vampomi_host.h"

int main(int argc, char** argv) { return vampomi_main(argc, argv); }
