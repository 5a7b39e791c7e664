/* oracle/refshim/opencv2/core.hpp -- TEST INFRASTRUCTURE ONLY: the name include/img_completion.h looks for; see opencv.hpp */
#include "opencv.hpp"
