#pragma once
#include "caffe2/core/operator.h"
