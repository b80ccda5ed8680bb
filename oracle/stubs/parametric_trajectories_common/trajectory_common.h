// TEST INFRASTRUCTURE.  Stand-in for the reference's PRIVATE dependency
// `parametric_trajectories_common::TPath` (not in /root/reference): the seven getters
// src/nmpc_nav_control/PathDiscretizer.cpp calls, over the curve families SURVEY.md 8(f2) names.
//
// One segment, parameter u in [0, 1]; sixteen doubles, the layout include/nmpc_b200.h documents as nmpc_path_segment:
//   kind 0  polynomial  x(u) = sum cx[i] u^i, y(u) = sum cy[i] u^i, i = 0..5 (Horner)   - lines, cubic Beziers, quintics
//   kind 1  arc         x(u) = cx[0] + cx[1] cos(cx[2] + cx[3] u), y(u) = cy[0] + cx[1] sin(cx[2] + cx[3] u)
//   vel     signed speed along the segment (negative = driven backwards)
//   th0,th1 holonomic heading at u = 0 and u = 1 (linear in between)
// GetTheta is the tangent direction atan2(dy/du, dx/du).
#pragma once
#include <cmath>

namespace parametric_trajectories_common {

class TPath {
  public:
    double kind = 0.0, vel = 0.0, th0 = 0.0, th1 = 0.0, cx[6] = {0, 0, 0, 0, 0, 0}, cy[6] = {0, 0, 0, 0, 0, 0};

    double GetVelocity() const { return vel; }
    double GetX(double u) const { return kind == 0.0 ? poly(cx, u) : cx[0] + cx[1] * std::cos(cx[2] + cx[3] * u); }
    double GetY(double u) const { return kind == 0.0 ? poly(cy, u) : cy[0] + cx[1] * std::sin(cx[2] + cx[3] * u); }
    double GetDX(double u) const { return kind == 0.0 ? dpoly(cx, u) : -cx[1] * cx[3] * std::sin(cx[2] + cx[3] * u); }
    double GetDY(double u) const { return kind == 0.0 ? dpoly(cy, u) : cx[1] * cx[3] * std::cos(cx[2] + cx[3] * u); }
    double GetTheta(double u) const { return std::atan2(GetDY(u), GetDX(u)); }
    double GetThetaHolomonic(double u) const { return th0 + (th1 - th0) * u; }

  private:
    static double poly(const double* c, double u) { return ((((c[5] * u + c[4]) * u + c[3]) * u + c[2]) * u + c[1]) * u + c[0]; }
    static double dpoly(const double* c, double u)
    {
        return (((5.0 * c[5] * u + 4.0 * c[4]) * u + 3.0 * c[3]) * u + 2.0 * c[2]) * u + c[1];
    }
};

}  // namespace parametric_trajectories_common
