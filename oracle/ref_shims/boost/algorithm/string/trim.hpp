// TEST INFRASTRUCTURE ONLY — the reference includes this header (options.cpp:9) but uses nothing from it.
#pragma once
