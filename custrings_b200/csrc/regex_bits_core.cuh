#pragma once
#include "regex_bits.h"
