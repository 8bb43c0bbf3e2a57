// ctopprm_main.cpp -- the reference's ./main flow (src/main.cpp:203-308) on the accelerated classes:
// YAML config -> ESDFMap::load -> TopologicalPRMClustering::find_geometrical_paths -> the same CSV
// outputs (<out><config>method_clustering_{time,num,length}.csv, appended one number per line with
// std::to_string; <out>roadmap_clustering_roadmap_shortened_unique_path0_<k>.csv with 6-column rows;
// <out>roadmap_all0.csv written by the planner).  Command line: the reference's TCLAP options that
// matter for the path (--config replaces the hard-coded "./config_files/building.yaml" of
// src/main.cpp:30; --output_folder, --map, --start_p x y z, --goal_p x y z as src/main.cpp:59-122).
//
//   g++ -std=c++17 -Iinclude examples/ctopprm_main.cpp -Lctopprm_learn_b200/lib -lctopprm_host \
//       -lctopprm_b200 -Wl,-rpath,$PWD/ctopprm_learn_b200/lib -o ctopprm_main
#include <chrono>
#include <cstring>
#include <fstream>
#include <memory>

#include "ctopprm/esdf_map.hpp"
#include "ctopprm/topological_prm_clustering.hpp"

static void save_double(const std::string &filename, double d) {  // src/main.cpp:156-165
  std::ofstream f(filename.c_str(), std::ios_base::app);
  if (f.is_open()) f << std::to_string(d) << std::endl;
}

int main(int argc, char **argv) {
  std::string planner_config_file = "./config_files/building.yaml", output_folder = "./prints/", map_override;
  Vector<3> start, goal;
  bool start_set = false, goal_set = false;
  for (int i = 1; i < argc; i++) {
    auto need = [&](int k) {
      if (i + k >= argc) {
        ERROR("missing value for " << argv[i]);
        exit(1);
      }
    };
    if (!strcmp(argv[i], "--config")) { need(1); planner_config_file = argv[++i]; }
    else if (!strcmp(argv[i], "--output_folder")) { need(1); output_folder = argv[++i]; }
    else if (!strcmp(argv[i], "--map")) { need(1); map_override = argv[++i]; }
    else if (!strcmp(argv[i], "--start_p")) { need(3); start = Vector<3>(atof(argv[i + 1]), atof(argv[i + 2]), atof(argv[i + 3])); i += 3; start_set = true; }
    else if (!strcmp(argv[i], "--goal_p")) { need(3); goal = Vector<3>(atof(argv[i + 1]), atof(argv[i + 2]), atof(argv[i + 3])); i += 3; goal_set = true; }
    else { ERROR("unknown option " << argv[i]); return 1; }
  }
  YAML::Node planner_config = YAML::LoadFile(planner_config_file);
  const std::string map_type = loadParam<std::string>(planner_config, "map_type");
  std::string map_file = loadParam<std::string>(planner_config, "map");
  if (!map_override.empty()) map_file = map_override;
  loadParam<double>(planner_config, "min_clearance");
  loadParam<bool>(planner_config, "check_collisions");
  loadParam<double>(planner_config, "collision_distance_check");
  if (!start_set && !parseArrayParam(planner_config["start"], "position", start)) {
    INFO_RED("you must specify start position");
    return 1;
  }
  if (!goal_set && !parseArrayParam(planner_config["end"], "position", goal)) {
    INFO_RED("you must specify end position");
    return 1;
  }
  std::shared_ptr<BaseMap> map;
  if (map_type == "ESDF") map = std::make_shared<ESDFMap>();
  else {
    ERROR("map type " << map_type << " not recognized");
    return 1;
  }
  map->load(map_file);
  std::vector<Vector<3>> gates = {start, goal};
  const std::string tag = output_folder + planner_config_file;  // file names as src/main.cpp:269-271 builds them
  auto t0 = std::chrono::high_resolution_clock::now();
  auto paths = TopologicalPRMClustering<Vector<3>>::find_geometrical_paths(planner_config, map, gates, output_folder);
  auto t1 = std::chrono::high_resolution_clock::now();
  const double secs = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count() * 1e-9;
  const size_t found = paths.empty() ? 0 : paths[0].size();
  std::cout << "number of paths found " << found << std::endl;
  save_double(tag + "method_clustering_time.csv", secs);
  save_double(tag + "method_clustering_num.csv", (double)found);
  for (size_t i = 0; i < found; i++) {
    TopologicalPRMClustering<Vector<3>>::savePath(
        output_folder + "roadmap_clustering_roadmap_shortened_unique_path0_" + std::to_string(i) + ".csv", paths[0][i].plan);
    save_double(tag + "method_clustering_length.csv", paths[0][i].length);
  }
  return 0;
}
