// On-disk formats of the reference that sit either side of the hot path, host code only:
//   * trajectory file   my_slam::vo::writePoseToFile / readPoseFromFile (reference src/vo/vo_io.cpp:51-120): one pose per
//                       line, 12 numbers "tx ty tz R00 R10 R20 R01 R11 R21 R02 R12 R22" (translation, then the rotation
//                       column by column), written with C++ stream defaults (6 significant digits), read as a flat
//                       whitespace-separated stream of doubles
//   * image names       readImagePaths (vo_io.cpp:12-37): dataset_dir + boost::format("/rgb_%05d.png") % i (run_vo.cpp:90)
//   * config/config.yaml in OpenCV's "%YAML:1.0" dialect, read by basics::Config / basics::Yaml: flat `key: value`
//                       scalars, one level of nesting for the dataset sections (`matlab:` / `fr1_desk:` with keys such as
//                       `camera_info.fx`), strings optionally double-quoted, `#` comments.  mvo_config_apply maps the
//                       keys the hot path latches (feature_match.cpp:16-23,56-58,137-139; vo.cpp / vo.h tracking keys)
//                       onto mvo_params / mvo_track_params with the reference's conversions (FileNode -> int rounds).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include "mvo.h"

struct mvo_config {
  std::map<std::string, std::string> kv;      // "key" or "section/key" -> raw scalar text (quotes stripped)
  std::string err;
};

namespace {

std::string trim(const std::string &s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return s.substr(a, b - a);
}

// value text up to an unquoted '#', quotes removed
std::string scalar_of(const std::string &raw) {
  std::string v;
  bool in_q = false;
  for (char c : raw) {
    if (c == '"') { in_q = !in_q; continue; }
    if (c == '#' && !in_q) break;
    v.push_back(c);
  }
  return trim(v);
}

bool get_raw(const mvo_config *c, const char *key, std::string *out) {
  auto it = c->kv.find(key);
  if (it == c->kv.end()) return false;
  *out = it->second;
  return true;
}

}  // namespace

extern "C" {

int mvo_write_pose_file(const char *filename, const double *T_list, int n) {
  if (!filename || (n > 0 && !T_list) || n < 0) return MVO_ERR_INVALID_ARG;
  std::ofstream fout(filename);
  if (!fout.is_open()) return MVO_ERR_INVALID_ARG;                  // the reference prints a warning and returns (:55-60)
  for (int k = 0; k < n; ++k) {
    const double *T = T_list + 16 * (size_t)k;
    fout << T[3] << " " << T[7] << " " << T[11] << " ";
    for (int i = 0; i < 3; ++i)                                     // order: 1st column, 2nd column, 3rd column (:67-73)
      for (int j = 0; j < 3; ++j) fout << T[j * 4 + i] << " ";
    fout << '\n';
  }
  return fout.good() ? MVO_OK : MVO_ERR_INVALID_ARG;
}

int mvo_read_pose_file(const char *filename, double *T_list, int cap, int *n_out) {
  if (!filename || !n_out || cap < 0 || (cap > 0 && !T_list)) return MVO_ERR_INVALID_ARG;
  *n_out = 0;
  std::ifstream fin(filename);
  if (!fin.is_open()) return MVO_ERR_INVALID_ARG;                   // the reference asserts (:89)
  double pose[12], val;
  int cnt = 0, n = 0;
  while (fin >> val) {
    pose[cnt++] = val;
    if (cnt == 12) {                                                // :97-112
      cnt = 0;
      if (n < cap) {
        double *T = T_list + 16 * (size_t)n;
        T[0] = pose[3]; T[1] = pose[6]; T[2] = pose[9];   T[3] = pose[0];
        T[4] = pose[4]; T[5] = pose[7]; T[6] = pose[10];  T[7] = pose[1];
        T[8] = pose[5]; T[9] = pose[8]; T[10] = pose[11]; T[11] = pose[2];
        T[12] = T[13] = T[14] = 0; T[15] = 1;
      }
      ++n;
    }
  }
  *n_out = n;
  return n > cap ? MVO_ERR_CAPACITY : MVO_OK;
}

int mvo_image_path(const char *dataset_dir, const char *image_formatting, int index, char *out, size_t cap) {
  if (!dataset_dir || !image_formatting || !out || cap == 0) return MVO_ERR_INVALID_ARG;
  // boost::format(dataset_dir + image_formatting) % index with a single %0Nd directive (run_vo.cpp:90: "/rgb_%05d.png")
  const std::string fmt = std::string(dataset_dir) + image_formatting;
  const size_t p = fmt.find('%');
  if (p == std::string::npos) return MVO_ERR_INVALID_ARG;
  size_t q = p + 1;
  while (q < fmt.size() && (fmt[q] >= '0' && fmt[q] <= '9')) ++q;
  if (q >= fmt.size() || (fmt[q] != 'd' && fmt[q] != 'i')) return MVO_ERR_INVALID_ARG;
  const std::string spec = fmt.substr(p, q - p) + "d";
  char num[64];
  snprintf(num, sizeof num, spec.c_str(), index);
  const std::string res = fmt.substr(0, p) + num + fmt.substr(q + 1);
  if (res.size() + 1 > cap) return MVO_ERR_CAPACITY;
  memcpy(out, res.c_str(), res.size() + 1);
  return MVO_OK;
}

int mvo_config_load(const char *filename, mvo_config **out) {
  if (!filename || !out) return MVO_ERR_INVALID_ARG;
  *out = nullptr;
  std::ifstream fin(filename);
  if (!fin.is_open()) return MVO_ERR_INVALID_ARG;                   // config.cpp:19-23: "Parameter file ... does not exist."
  mvo_config *c = new mvo_config();
  std::string line, section;
  while (std::getline(fin, line)) {
    if (line.rfind("%YAML", 0) == 0 || line.rfind("---", 0) == 0) continue;
    const size_t first = line.find_first_not_of(" \t\r");
    if (first == std::string::npos || line[first] == '#') continue;
    const size_t colon = line.find(':', first);
    if (colon == std::string::npos) continue;
    const std::string key = trim(line.substr(first, colon - first));
    const std::string val = scalar_of(line.substr(colon + 1));
    if (first == 0) {
      section.clear();
      if (val.empty()) { section = key; continue; }                 // a nested map starts (dataset sections)
      c->kv[key] = val;
    } else {
      if (val.empty()) continue;
      c->kv[section.empty() ? key : section + "/" + key] = val;
    }
  }
  *out = c;
  return MVO_OK;
}

void mvo_config_free(mvo_config *c) { delete c; }

int mvo_config_get_string(const mvo_config *c, const char *key, char *out, size_t cap) {
  if (!c || !key || !out || cap == 0) return MVO_ERR_INVALID_ARG;
  std::string v;
  if (!get_raw(c, key, &v)) return MVO_ERR_INVALID_ARG;             // config.cpp:35 throws "Key ... doesn't exist"
  if (v.size() + 1 > cap) return MVO_ERR_CAPACITY;
  memcpy(out, v.c_str(), v.size() + 1);
  return MVO_OK;
}

int mvo_config_get_double(const mvo_config *c, const char *key, double *out) {
  if (!c || !key || !out) return MVO_ERR_INVALID_ARG;
  std::string v;
  if (!get_raw(c, key, &v)) return MVO_ERR_INVALID_ARG;
  char *end = nullptr;
  const double d = strtod(v.c_str(), &end);
  if (end == v.c_str()) return MVO_ERR_INVALID_ARG;
  *out = d;
  return MVO_OK;
}

int mvo_config_get_int(const mvo_config *c, const char *key, int *out) {
  double d;
  const int rc = mvo_config_get_double(c, key, &d);
  if (rc != MVO_OK) return rc;
  *out = (int)std::nearbyint(d);                                    // cv::FileNode -> int is cvRound
  return MVO_OK;
}

int mvo_config_get_bool(const mvo_config *c, const char *key, int *out) {
  if (!c || !key || !out) return MVO_ERR_INVALID_ARG;
  std::string v;
  if (!get_raw(c, key, &v)) return MVO_ERR_INVALID_ARG;
  *out = (v == "true" || v == "True") ? 1 : 0;                      // config.cpp:41-47
  return MVO_OK;
}

int mvo_config_apply(const mvo_config *c, mvo_params *p, mvo_track_params *tp, double *K9) {
  if (!c) return MVO_ERR_INVALID_ARG;
  int rc = MVO_OK, iv;
  double dv;
#define GET_I(key, dst) do { if ((rc = mvo_config_get_int(c, key, &iv)) != MVO_OK) return rc; dst = iv; } while (0)
#define GET_D(key, dst) do { if ((rc = mvo_config_get_double(c, key, &dv)) != MVO_OK) return rc; dst = dv; } while (0)
  if (p) {
    GET_I("number_of_keypoints_to_extract", p->orb_nfeatures);      // feature_match.cpp:16-21
    GET_D("scale_factor", p->orb_scale_factor);
    GET_I("level_pyramid", p->orb_nlevels);
    GET_I("score_threshold", p->orb_fast_threshold);
    GET_I("max_number_of_keypoints", p->max_keypoints);             // :56-58
    GET_I("kpts_uniform_selection_grid_size", p->grid_size);
    GET_I("kpts_uniform_selection_max_pts_per_grid", p->max_pts_per_grid);
    GET_I("xiang_gao_method_match_ratio", p->xiang_gao_ratio);      // :137-139, read with get<int> in the reference
    GET_I("lowe_method_dist_ratio", p->lowe_ratio);
  }
  if (tp) {
    GET_I("feature_match_method_index_pnp", tp->match_method);      // vo.cpp:283-289
    GET_D("max_matching_pixel_dist_in_pnp", tp->match_radius);
    GET_D("max_possible_dist_to_prev_keyframe", tp->max_dist_to_prev);
    GET_D("min_dist_between_two_keyframes", tp->min_dist_keyframe);
    GET_I("num_prev_frames_to_opti_by_ba", tp->ba_window);
    if ((rc = mvo_config_get_bool(c, "is_enable_ba", &iv)) != MVO_OK) return rc;
    tp->ba_enable = iv;
    if ((rc = mvo_config_get_bool(c, "is_ba_fix_map_points", &iv)) != MVO_OK) return rc;
    tp->ba_fix_points = iv;
    std::string info;
    if (!get_raw(c, "information_matrix", &info)) return MVO_ERR_INVALID_ARG;
    std::istringstream is(info);                                    // "1.0 0.0 0.0 1.0" (vo.cpp:406-413)
    for (int i = 0; i < 4; ++i) if (!(is >> tp->information[i])) return MVO_ERR_INVALID_ARG;
  }
  if (K9) {                                                         // readCameraIntrinsics (vo_io.cpp:40-49) of the selected dataset
    std::string ds;
    if (!get_raw(c, "dataset_name", &ds)) return MVO_ERR_INVALID_ARG;
    double fx, fy, cx, cy;
    GET_D((ds + "/camera_info.fx").c_str(), fx);
    GET_D((ds + "/camera_info.fy").c_str(), fy);
    GET_D((ds + "/camera_info.cx").c_str(), cx);
    GET_D((ds + "/camera_info.cy").c_str(), cy);
    const double K[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1};
    memcpy(K9, K, sizeof K);
  }
#undef GET_I
#undef GET_D
  return MVO_OK;
}

int mvo_config_apply_vo(const mvo_config *c, mvo_vo_params *vp) {
  if (!c || !vp) return MVO_ERR_INVALID_ARG;
  int rc = mvo_config_apply(c, nullptr, &vp->track, nullptr), iv;
  if (rc != MVO_OK) return rc;
  double dv;
#define GET_I(key, dst) do { if ((rc = mvo_config_get_int(c, key, &iv)) != MVO_OK) return rc; dst = iv; } while (0)
#define GET_D(key, dst) do { if ((rc = mvo_config_get_double(c, key, &dv)) != MVO_OK) return rc; dst = dv; } while (0)
  GET_I("feature_match_method_index_initialization", vp->match_method_init);          // vo_addFrame.cpp:41
  GET_D("max_matching_pixel_dist_in_initialization", vp->max_match_dist_init);        // :39-40
  GET_D("max_matching_pixel_dist_in_triangulation", vp->max_match_dist_triangulation);   // :97-98
  GET_D("findEssentialMat_threshold", vp->essential_threshold);                       // epipolar_geometry.cpp:32
  GET_D("min_triang_angle", vp->min_triang_angle);                                    // vo.cpp:183-185
  GET_D("max_ratio_between_max_angle_and_median_angle", vp->max_ratio_angle_to_median);
  GET_I("min_inlier_matches", vp->min_inlier_matches);                                // vo.cpp:123-125
  GET_D("min_pixel_dist", vp->min_pixel_dist);
  GET_D("min_median_triangulation_angle", vp->min_median_triangulation_angle);
  GET_D("assumed_mean_pts_depth_during_vo_init", vp->assumed_mean_depth_init);        // vo.cpp:103-104
#undef GET_I
#undef GET_D
  return MVO_OK;
}

}  // extern "C"
