// main_meth_probit — the entry point BASELINE.json's configuration 4 names (reference: src/main_meth_probit.cpp, which does not
// compile against the shipped vamp.hpp): main_meth with the probit model forced, plus that driver's `test` and `predict` run modes.
#include "../../../include/vampomi_host.h"

int main(int argc, char** argv) { return vampomi_main_probit(argc, argv); }
