// eigen_facade: units.h:7 includes <tbb/tbb.h>; the only tbb:: use in the headers compiled here sits under #ifdef NEW_CODE (util.h:158-187), which is off.
// The real header drags in the standard headers pcg.h relies on without including them itself.
#pragma once
#include <chrono>
#include <cmath>
#include <iostream>
