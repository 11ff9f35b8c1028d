#pragma once
#include "parallel_for.h"
