#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using std::filesystem::path; using std::filesystem::exists; } }
