// B200 facade of macGrid/macGridCell.h:8-58.  On the device a cell is spread over fp32 SoA channels; this struct is the
// HOST MIRROR element that MacGrid::cell() hands out (filled lazily by fsim_download_grid).  Member names and types follow
// the reference (std::atomic<double> included) because SimulationManager and the GUI inspector read them directly.
#pragma once
#include <atomic>
#include <glm/glm.hpp>

namespace genericfsim::macgrid {

struct MacGridCell {
    struct Face {
        std::atomic<double> v{0.0};
        double v2 = 0;
        std::atomic<double> particleWeightSum{0.0};
        glm::dvec3 pos;

        Face() = default;
        explicit Face(const glm::dvec3& p) : pos(p) {}
        Face(const Face& o) : v(o.v.load()), v2(o.v2), particleWeightSum(o.particleWeightSum.load()), pos(o.pos) {}
        Face& operator=(const Face& o) {
            v = o.v.load(); v2 = o.v2; particleWeightSum = o.particleWeightSum.load(); pos = o.pos;
            return *this;
        }
    };
    struct FaceRef { Face& face; };
    enum class CellType { WATER, AIR, SOLID };  // same order as FSIM_CELL_*

    Face faces[3];
    glm::dvec3 pos;
    CellType type = CellType::AIR;
    std::atomic<double> avgPNum{0.0};
    int id = 0;

    MacGridCell() = default;
    MacGridCell(const MacGridCell& o) : faces{o.faces[0], o.faces[1], o.faces[2]}, pos(o.pos), type(o.type), avgPNum(o.avgPNum.load()), id(o.id) {}
    MacGridCell& operator=(const MacGridCell& o) {
        for (int a = 0; a < 3; a++) faces[a] = o.faces[a];
        pos = o.pos; type = o.type; avgPNum = o.avgPNum.load(); id = o.id;
        return *this;
    }
};

struct MacGridCellRef { MacGridCell& cell; };

}  // namespace genericfsim::macgrid
