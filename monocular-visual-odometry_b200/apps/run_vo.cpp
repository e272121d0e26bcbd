// run_vo — the reference's application (reference run_vo.cpp:61-151) on libmvo, without the two display windows:
//   run_vo <config.yaml> [max_frames]
// reads the reference's config file (dataset_name, <dataset>/dataset_dir, num_images, camera_info.*, max_num_imgs_to_proc,
// save_predicted_traj_to and every algorithm key), the images dataset_dir + "/rgb_%05d.png" (run_vo.cpp:90), feeds them
// through the VisualOdometry state machine (mvo_vo_*) and writes the camera trajectory in writePoseToFile's format.
// One line per frame goes to stdout.  Exit code 0 on success, 1 on any error (no GPU, unreadable config ...).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "mvo.h"
#include "png_reader.h"

namespace {
struct Fail { std::string msg; };
void chk(int rc, mvo_ctx *ctx, const char *what) {
  if (rc == MVO_OK) return;
  throw Fail{std::string(what) + " failed (" + std::to_string(rc) + ")" + (ctx ? std::string(": ") + mvo_last_error(ctx) : std::string())};
}
std::string cfg_string(const mvo_config *c, const std::string &key) {
  char buf[1024];
  if (mvo_config_get_string(c, key.c_str(), buf, sizeof buf) != MVO_OK) throw Fail{"config key " + key + " is missing"};
  return buf;
}
int cfg_int(const mvo_config *c, const std::string &key) {
  int v = 0;
  if (mvo_config_get_int(c, key.c_str(), &v) != MVO_OK) throw Fail{"config key " + key + " is missing"};
  return v;
}
}  // namespace

int main(int argc, char **argv) {
  if (argc < 2) {                                                 // checkInputArguments (run_vo.cpp:160-170)
    fprintf(stderr, "Lack arguments: Please input the path to the .yaml config file\n");
    return 1;
  }
  mvo_config *cfg = nullptr;
  mvo_ctx *ctx = nullptr;
  mvo_vo *vo = nullptr;
  int status = 0;
  try {
    if (mvo_config_load(argv[1], &cfg) != MVO_OK) throw Fail{std::string("cannot read the config file ") + argv[1]};
    const std::string dataset = cfg_string(cfg, "dataset_name");
    const std::string dataset_dir = cfg_string(cfg, dataset + "/dataset_dir");
    const int num_images = cfg_int(cfg, dataset + "/num_images");
    int max_frames = cfg_int(cfg, "max_num_imgs_to_proc");
    if (argc > 2) max_frames = std::min(max_frames, atoi(argv[2]));
    const std::string traj_file = cfg_string(cfg, "save_predicted_traj_to");
    mvo_params p;
    mvo_vo_params vp;
    double K[9];
    mvo_default_params(&p);
    mvo_vo_default_params(&vp);
    chk(mvo_config_apply(cfg, &p, nullptr, K), nullptr, "config (ORB / matching keys, camera_info)");
    chk(mvo_config_apply_vo(cfg, &vp), nullptr, "config (VO keys)");
    const int rc = mvo_create(&ctx, 0, &p);
    if (rc != MVO_OK) throw Fail{"no usable sm_100 GPU (mvo_create returned " + std::to_string(rc) + "); libmvo has no CPU path"};
    std::vector<double> history;
    std::vector<uint8_t> bgr;
    const int n = std::min(max_frames, num_images);
    int rows0 = 0, cols0 = 0;
    for (int img_id = 0; img_id < n; ++img_id) {
      char path[2048];
      chk(mvo_image_path(dataset_dir.c_str(), "/rgb_%05d.png", img_id, path, sizeof path), ctx, "image path");
      int rows = 0, cols = 0;
      std::string err;
      if (!mvo_app::read_png_bgr(path, &bgr, &rows, &cols, &err)) {     // run_vo.cpp:115-119: an unreadable image ends the run
        printf("The image file %s is empty. Finished. (%s)\n", path, err.c_str());
        break;
      }
      if (!vo) { chk(mvo_vo_create(ctx, K, rows, cols, &vp, &vo), ctx, "mvo_vo_create"); rows0 = rows; cols0 = cols; }
      if (rows != rows0 || cols != cols0)
        throw Fail{std::string(path) + ": " + std::to_string(cols) + "x" + std::to_string(rows) + " differs from the first frame's " + std::to_string(cols0) + "x" + std::to_string(rows0)};
      double T[16];
      mvo_vo_frame_info info;
      chk(mvo_vo_add_frame(vo, bgr.data(), 3, (size_t)cols * 3, T, &info), ctx, "addFrame");
      history.insert(history.end(), T, T + 16);
      printf("frame %d: state %d -> %d%s, %d keypoints, %d matches, %d inliers, map %d, t = %.5f %.5f %.5f\n", info.frame_id, info.state_in,
             info.state_out, info.keyframe ? " keyframe" : "", info.n_keypoints, info.n_matches, info.n_inliers, info.map_points, T[3], T[7], T[11]);
    }
    chk(mvo_write_pose_file(traj_file.c_str(), history.data(), (int)(history.size() / 16)), nullptr, "writing the trajectory");
    printf("Wrote %d poses to %s\n", (int)(history.size() / 16), traj_file.c_str());
  } catch (const Fail &f) {
    fprintf(stderr, "run_vo: %s\n", f.msg.c_str());
    status = 1;
  } catch (const std::exception &e) {                            // e.g. std::bad_alloc from a huge image
    fprintf(stderr, "run_vo: %s\n", e.what());
    status = 1;
  }
  mvo_vo_destroy(vo);
  if (ctx) mvo_destroy(ctx);
  if (cfg) mvo_config_free(cfg);
  return status;
}
