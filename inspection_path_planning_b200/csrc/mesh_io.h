// mesh_io.h -- mesh ingest for libipp.so (see mesh_io.cpp)
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

namespace ipp {
// Loads .stl / .obj / .ippm; `tris` receives 9 floats per triangle in file order.
// Throws std::invalid_argument like SceneKeeper::LoadScene (SceneKeeper.cpp:79-82).
void load_mesh(const std::string& path, std::vector<float>& tris);
}  // namespace ipp
