#pragma once
#include <PapillonNDL/st_neutron.hpp>
