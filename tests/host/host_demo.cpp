// Exercises the drop-in C++ classes exactly the way Frame::ExtractORB / Frame::ExtractLSD call the reference's
// (Frame.cc:322-334): reads a raw 8-bit frame, writes keypoints/descriptors/keylines to a binary file for comparison.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../pl-slam_b200/host/ORBextractor.h"
#include "../../pl-slam_b200/host/LineExtractor.h"
int main(int argc, char** argv) {
  if (argc < 5) { fprintf(stderr, "usage: host_demo frame.raw W H out.bin\n"); return 2; }
  const int W = atoi(argv[2]), H = atoi(argv[3]);
  std::vector<uint8_t> buf((size_t)W * H);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(buf.data(), 1, buf.size(), f) != buf.size()) { fprintf(stderr, "cannot read frame\n"); return 2; }
  fclose(f);
  cv::Mat im(H, W, CV_8UC1, buf.data()), none;
  ORB_SLAM2::ORBextractor orb(1000, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> keys; cv::Mat desc;
  orb(im, none, keys, desc);
  ORB_SLAM2::LINEextractor lsd(1, 1.2f, 200, 0.0);
  std::vector<KeyLine> lines; cv::Mat ldesc; std::vector<Eigen::Vector3d> lf;
  lsd(im, none, lines, ldesc, lf);
  FILE* o = fopen(argv[4], "wb");
  int n = (int)keys.size(), nl = (int)lines.size();
  fwrite(&n, 4, 1, o); fwrite(keys.data(), sizeof(cv::KeyPoint), n, o); for (int i = 0; i < n; i++) fwrite(desc.ptr(i), 1, 32, o);
  fwrite(&nl, 4, 1, o); fwrite(lines.data(), sizeof(KeyLine), nl, o); for (int i = 0; i < nl; i++) fwrite(ldesc.ptr(i), 1, 32, o);
  fclose(o);
  printf("host_demo: %d keypoints, %d keylines (levels %d, scale %.2f)\n", n, nl, orb.GetLevels(), orb.GetScaleFactor());
  return 0;
}
