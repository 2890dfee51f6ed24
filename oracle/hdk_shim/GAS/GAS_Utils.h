// hdk_shim: stands in for <GAS/GAS_Utils.h> of the Houdini Development Kit (absent offline); everything lives in hdk_shim.h
#pragma once
#include "../hdk_shim.h"
