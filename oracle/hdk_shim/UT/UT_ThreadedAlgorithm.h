// hdk_shim: stands in for <UT/UT_ThreadedAlgorithm.h> of the Houdini Development Kit (absent offline); everything lives in hdk_shim.h
#pragma once
#include "../hdk_shim.h"
