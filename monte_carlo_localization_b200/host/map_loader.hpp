// host/map_loader.hpp -- maps/<name>.yaml + image -> occupancy grid, without ROS.
//
// In the reference the grid arrives from nav2_map_server over the /map_server/map service
// (launch/mcl_launch.py:62-71, src/particle_filter.cpp:148, 184-190).  nav2_map_server is not
// vendored, so its published trinary conversion is restated here (and in maps.py, which the
// tests cross-check against this file):
//   shade = mean(r, g, b [, alpha for images with alpha]) / 255
//   occ   = negate ? shade : 1 - shade
//   cell  = occ > occupied_thresh ? 100 : occ < free_thresh ? 0 : -1
//   grid row 0 = bottom row of the image.
// Supported images: PGM (P5, P2) and non-interlaced 8-bit PNG (gray, gray+alpha, RGB, RGBA,
// palette) -- every image shipped under the reference's maps/.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace particle_filter_cpp {

struct OccupancyGrid {          // the fields of nav_msgs/OccupancyGrid that get_omap reads
    std::vector<int8_t> data;   // row-major, row 0 = bottom
    uint32_t width = 0, height = 0;
    float resolution = 0.f;     // float32, as in the message
    double origin_x = 0, origin_y = 0, origin_yaw = 0;
};

struct MapYaml {
    std::string image;
    double resolution = 0, occupied_thresh = 0.65, free_thresh = 0.196;
    double origin[3] = {0, 0, 0};
    int negate = 0;
};

struct Image8 {
    int width = 0, height = 0, channels = 0;   // 1 gray, 2 gray+alpha, 3 rgb, 4 rgba
    std::vector<uint8_t> pix;                   // row-major, top row first
};

bool parse_map_yaml(const std::string& path, MapYaml& out, std::string* err);
bool load_image(const std::string& path, Image8& out, std::string* err);
void image_to_grid(const Image8& img, const MapYaml& meta, OccupancyGrid& out);
// yaml + image -> grid, as the map server would serve it
bool load_map(const std::string& yaml_path, OccupancyGrid& out, std::string* err);

}  // namespace particle_filter_cpp
