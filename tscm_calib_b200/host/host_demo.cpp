// host_demo.cpp — exercises the drop-in C++ adapters end to end (used by tests/ on the GPU
// box): reads a problem dumped by tests (binary, little-endian), builds a MultiCalib (or a
// TripleSphereCamera for C == 1), calibrates on the B200 through tscm_solve(), writes the
// YAML and a binary result the test compares with the oracle.
//
//   file layout: int32 C, F, K; double board_xy[K*2]; uint8 visible[C*F];
//                per camera m, per frame i visible: double obs[K*2];
//                double intr[C*9], cam_rt[C*6], board_rt[F*6]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "multi_calib_b200.h"

static bool read_all(FILE* f, void* p, size_t n) { return std::fread(p, 1, n, f) == n; }

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: host_demo problem.bin result.bin calib.yaml [--yaml-only] [--gpus N]\n"); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) return 2;
  int32_t C, F, K;
  read_all(f, &C, 4); read_all(f, &F, 4); read_all(f, &K, 4);
  std::vector<double> board(2 * (size_t)K);
  read_all(f, board.data(), board.size() * 8);
  std::vector<uint8_t> vis((size_t)C * F);
  read_all(f, vis.data(), vis.size());
  std::vector<std::vector<std::vector<cv::Point2d>>> pix(C, std::vector<std::vector<cv::Point2d>>(F));
  for (int m = 0; m < C; ++m)
    for (int i = 0; i < F; ++i)
      if (vis[(size_t)m * F + i]) {
        pix[m][i].resize(K);
        read_all(f, pix[m][i].data(), sizeof(double) * 2 * K);
      }
  std::vector<double> intr(9 * (size_t)C), cam_rt(6 * (size_t)C), board_rt(6 * (size_t)F);
  read_all(f, intr.data(), intr.size() * 8);
  read_all(f, cam_rt.data(), cam_rt.size() * 8);
  read_all(f, board_rt.data(), board_rt.size() * 8);
  std::fclose(f);

  std::vector<cv::Point3d> worlds(K);
  for (int j = 0; j < K; ++j) worlds[j] = cv::Point3d(board[2 * j], board[2 * j + 1], 0.0);
  auto rot = [](const double* r) { cv::Mat v(3, 1), R; for (int k = 0; k < 3; ++k) v.at<double>(k, 0) = r[k]; cv::Rodrigues(v, R); return R; };
  auto tr = [](const double* t) { cv::Mat v(3, 1); for (int k = 0; k < 3; ++k) v.at<double>(k, 0) = t[k]; return v; };

  std::vector<MultiCalib_camera> cams;
  for (int m = 0; m < C; ++m) {
    std::vector<bool> has(F);
    for (int i = 0; i < F; ++i) has[i] = vis[(size_t)m * F + i] != 0;
    const double* in = &intr[9 * m];
    MultiCalib_camera cam(in[2], in[3], in[0], in[1], in[4], in[5], in[6], in[7], in[8], rot(&cam_rt[6 * m]),
                          tr(&cam_rt[6 * m + 3]), has, pix[m]);
    // keep the exact initial angle-axis values (Rodrigues round trips are not bit exact)
    std::memcpy(cam.rt_.data(), &cam_rt[6 * m], 48);
    cams.push_back(cam);
  }
  std::vector<MultiCalib_chessboard> boards;
  for (int i = 0; i < F; ++i) {
    MultiCalib_chessboard b(rot(&board_rt[6 * i]), tr(&board_rt[6 * i + 3]));
    std::memcpy(b.rt_.data(), &board_rt[6 * i], 48);
    boards.push_back(b);
  }
  MultiCalib calib(cams, boards, worlds);
  bool yaml_only = false;                                                       // no GPU needed
  for (int k = 4; k < argc; ++k) {
    if (std::string(argv[k]) == "--yaml-only") yaml_only = true;
    // frames sharded over N GPUs of this process (tscm_options.num_gpus)
    if (std::string(argv[k]) == "--gpus" && k + 1 < argc) calib.options().num_gpus = std::atoi(argv[++k]);
  }
  if (!yaml_only) calib.calibrate();
  if (!calib.write_yaml(argv[3])) return 3;

  FILE* o = std::fopen(argv[2], "wb");
  if (!o) return 3;
  const tscm_summary& s = calib.last_summary();
  int32_t head[4] = {s.termination_type, s.num_iterations, s.num_successful_steps, s.num_unsuccessful_steps};
  std::fwrite(head, 4, 4, o);
  double costs[3] = {s.initial_cost, s.final_cost, calib.average_reprojection_error};
  std::fwrite(costs, 8, 3, o);
  for (int m = 0; m < C; ++m) std::fwrite(calib.cameras_[m].intrinsic_.data(), 8, 9, o);
  for (int m = 0; m < C; ++m) std::fwrite(calib.cameras_[m].rt_.data(), 8, 6, o);
  for (int i = 0; i < F; ++i) std::fwrite(calib.chessboards_[i].rt_.data(), 8, 6, o);
  calib.camera_reprojection_error.resize(C, 0.0);
  std::fwrite(calib.camera_reprojection_error.data(), 8, C, o);
  std::fclose(o);
  return 0;
}
