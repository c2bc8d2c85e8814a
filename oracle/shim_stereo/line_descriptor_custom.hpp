// Stand-in for the reference's 3rdparty line_descriptor headers: the KeyLine record (same public fields as
// cv::line_descriptor::KeyLine) and declaration-level LSD / LBD classes (never called on the matching path).
// Test infrastructure only.
#pragma once
#include <opencv2/core.hpp>

namespace cv { namespace line_descriptor {

struct KeyLine {
    float angle = 0.f;
    int class_id = -1, octave = 0;
    Point2f pt;
    float response = 0.f, size = 0.f;
    float startPointX = 0.f, startPointY = 0.f, endPointX = 0.f, endPointY = 0.f;
    float sPointInOctaveX = 0.f, sPointInOctaveY = 0.f, ePointInOctaveX = 0.f, ePointInOctaveY = 0.f;
    float lineLength = 0.f;
    int numOfPixels = 0;
};

class BinaryDescriptor {
public:
    static Ptr<BinaryDescriptor> createBinaryDescriptor() {
        standin_unavailable("BinaryDescriptor");
        return Ptr<BinaryDescriptor>();
    }
    void compute(const Mat &, std::vector<KeyLine> &, Mat &) { standin_unavailable("BinaryDescriptor::compute"); }
};

class LSDDetectorC {
public:
    struct LSDOptions {
        int refine = 0, n_bins = 0;
        double scale = 0, sigma_scale = 0, quant = 0, ang_th = 0, log_eps = 0, density_th = 0, min_length = 0;
    };
    static Ptr<LSDDetectorC> createLSDDetectorC() {
        standin_unavailable("LSDDetectorC");
        return Ptr<LSDDetectorC>();
    }
    void detect(const Mat &, std::vector<KeyLine> &, int, int, LSDOptions) { standin_unavailable("LSDDetectorC::detect"); }
    void detect(const Mat &, std::vector<KeyLine> &, double, int, LSDOptions) { standin_unavailable("LSDDetectorC::detect"); }
};

}} // namespace cv::line_descriptor
namespace line_descriptor = cv::line_descriptor;
