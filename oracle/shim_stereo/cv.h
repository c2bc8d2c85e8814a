#include <opencv2/core.hpp>
#include <opencv2/features2d.hpp>
