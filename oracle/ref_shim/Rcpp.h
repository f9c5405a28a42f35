#include "RcppArmadillo.h"
