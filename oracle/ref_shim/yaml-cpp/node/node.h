#pragma once
#include <yaml-cpp/yaml.h>
