// TEST INFRASTRUCTURE ONLY — included by the reference (vamp_probit.cpp:16), nothing used.
#pragma once
