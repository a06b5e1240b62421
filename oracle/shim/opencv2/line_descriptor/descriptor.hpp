// ORACLE shim (test infrastructure).  The reference includes opencv_contrib's <opencv2/line_descriptor/descriptor.hpp>, which is
// not in /root/reference; its tree vendors the same module as Thirdparty/line_descriptor (class names with a trailing C for the
// detector).  This header maps the contrib name to the vendored header (-I.../Thirdparty/line_descriptor/include).
#pragma once
#include <opencv2/core/core.hpp>
#include "line_descriptor_custom.hpp"
namespace cv { namespace line_descriptor {
class LSDDetector : public LSDDetectorC {
 public:
  static Ptr<LSDDetector> createLSDDetector() { return Ptr<LSDDetector>(new LSDDetector()); }
};
} }
