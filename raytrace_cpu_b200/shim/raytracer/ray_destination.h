// ray_destination.h (kerrtrace shim) -- the reference's stop-surface classes (src/raytracer/ray_destination.h:42-204)
// with the same names, constructors and virtual interface, plus one hook: to_kt_stop() describes the surface to the
// GPU engine.  The three concrete classes of the reference are offloaded; a user-defined subclass keeps the default
// hook and is rejected at run_raytrace() (its virtual methods cannot be called from device code).
#ifndef RAY_DESTINATION_H_
#define RAY_DESTINATION_H_

#include <cmath>
#include <limits>
#include <stdexcept>
#include <typeinfo>

#include "kerrtrace.h"

template <typename T>
class RayDestination {
public:
    virtual ~RayDestination() = default;
    virtual bool reached(T r, T theta, T phi) const = 0;
    virtual bool reached(T r, T theta, T phi, T prev_theta) const { return reached(r, theta, phi); }
    virtual T step_limit(T r, T theta, T phi, T pr, T ptheta, T pphi) const { return std::numeric_limits<T>::max(); }
    virtual T velocity(T r, T theta, T phi) const { return -1; }
    // Circular-orbit 4-velocity of the material at (r, theta, phi) -- host-side helper with the reference's formula
    virtual void four_velocity(T r, T theta, T phi, T spin, T et[4]) const
    {
        T V = velocity(r, theta, phi);
        const T sn = sin(theta), cs = cos(theta);
        const T rho2 = r * r + (spin * cs) * (spin * cs);
        const T del = r * r - 2 * r + spin * spin;
        const T sig2 = (r * r + spin * spin) * (r * r + spin * spin) - spin * spin * del * sn * sn;
        const T e2nu = rho2 * del / sig2, e2psi = sig2 * sn * sn / rho2, omega = 2 * spin * r / sig2;
        if (V == -1) V = 1 / (spin + r * sqrt(r));
        const T gam = 1 / sqrt(1 - (V - omega) * (V - omega) * e2psi / e2nu);
        et[0] = gam / sqrt(e2nu); et[1] = 0; et[2] = 0; et[3] = gam * V / sqrt(e2nu);
    }
    // shim hook: kind < 0 means "cannot run on the device"
    virtual kt_stop to_kt_stop() const { kt_stop s = {}; s.kind = -1; return s; }
};

template <typename T>
class FlatDiscDestination : public RayDestination<T> {
    T theta_lim;
public:
    explicit FlatDiscDestination(T theta_lim = M_PI_2) : theta_lim(theta_lim) {}
    bool reached(T, T theta, T) const override
    {
        return theta_lim > 0 ? theta >= theta_lim : (theta_lim < 0 ? theta <= -theta_lim : false);
    }
    T step_limit(T, T theta, T, T, T ptheta, T) const override
    {
        if (theta_lim > 0 && ptheta > 0 && theta < theta_lim) return (theta_lim - theta) / ptheta;
        if (theta_lim < 0 && ptheta < 0 && theta > -theta_lim) return (-theta_lim - theta) / ptheta;
        return std::numeric_limits<T>::max();
    }
    kt_stop to_kt_stop() const override { kt_stop s = {}; s.kind = KT_STOP_FLATDISC; s.theta_lim = theta_lim; return s; }
};

template <typename T>
class DiscWithISCODestination : public RayDestination<T> {
    T theta_lim, r_isco, r_out;
    bool in_annulus(T r) const { return !(r < r_isco) && !(r_out > 0 && r > r_out); }
public:
    explicit DiscWithISCODestination(T r_isco, T r_out = -1, T theta_lim = M_PI_2) : theta_lim(theta_lim), r_isco(r_isco), r_out(r_out) {}
    bool reached(T r, T theta, T) const override
    {
        if (!in_annulus(r)) return false;
        return theta_lim > 0 ? theta >= theta_lim : (theta_lim < 0 ? theta <= -theta_lim : false);
    }
    bool reached(T r, T theta, T, T prev_theta) const override
    {
        if (!in_annulus(r) || theta_lim == 0) return false;
        const T tl = theta_lim > 0 ? theta_lim : -theta_lim;
        return (prev_theta < tl && theta >= tl) || (prev_theta > tl && theta <= tl);
    }
    T step_limit(T r, T theta, T, T, T ptheta, T) const override
    {
        if (!in_annulus(r)) return std::numeric_limits<T>::max();
        if (theta_lim > 0 && ptheta > 0 && theta < theta_lim) return (theta_lim - theta) / ptheta;
        if (theta_lim < 0 && ptheta < 0 && theta > -theta_lim) return (-theta_lim - theta) / ptheta;
        return std::numeric_limits<T>::max();
    }
    kt_stop to_kt_stop() const override
    {
        kt_stop s = {};
        s.kind = KT_STOP_ISCODISC; s.theta_lim = theta_lim; s.r_isco = r_isco; s.r_out = r_out;
        return s;
    }
};

template <typename T>
class FlatPlaneDestination : public RayDestination<T> {
public:
    T incl, phi0, z_s;
    FlatPlaneDestination(T incl, T phi0, T z_s) : incl(incl), phi0(phi0), z_s(z_s) {}
    T projection(T r, T theta, T phi) const { return r * (sin(theta) * sin(incl) * cos(phi - phi0) + cos(theta) * cos(incl)); }
    bool reached(T r, T theta, T phi) const override { return projection(r, theta, phi) <= -z_s; }
    void source_coords(T r, T theta, T phi, T& x_s, T& y_s) const
    {
        const T X = r * sin(theta) * cos(phi), Y = r * sin(theta) * sin(phi), Z = r * cos(theta);
        x_s = -X * sin(phi0) + Y * cos(phi0);
        y_s = -X * cos(incl) * cos(phi0) - Y * cos(incl) * sin(phi0) + Z * sin(incl);
    }
    kt_stop to_kt_stop() const override
    {
        kt_stop s = {};
        s.kind = KT_STOP_FLATPLANE; s.incl = incl; s.phi0 = phi0; s.z_s = z_s;
        return s;
    }
};

namespace kerrtrace_shim {
template <typename T>
kt_stop stop_of(const RayDestination<T>* dest)
{
    if (!dest) throw std::runtime_error("kerrtrace: null RayDestination");
    // The reference's concrete classes are not final: a user class derived from one of them may override reached(),
    // step_limit(), velocity() or four_velocity() and would inherit the base class's to_kt_stop().  Only objects whose dynamic
    // type is EXACTLY one of the three known classes are offloaded; everything else is refused loudly.
    const std::type_info& ti = typeid(*dest);
    const bool known = ti == typeid(FlatDiscDestination<T>) || ti == typeid(DiscWithISCODestination<T>) ||
                       ti == typeid(FlatPlaneDestination<T>);
    const kt_stop s = known ? dest->to_kt_stop() : kt_stop{-1};
    if (s.kind < 0)
        throw std::runtime_error("kerrtrace: user-defined RayDestination subclasses are not offloaded (only FlatDisc, "
                                 "DiscWithISCO and FlatPlane destinations run on the device)");
    return s;
}
}  // namespace kerrtrace_shim

#endif /* RAY_DESTINATION_H_ */
