#pragma once
#include "ref_stub_common.h"
