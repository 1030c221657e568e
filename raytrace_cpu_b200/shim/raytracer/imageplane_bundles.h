// imageplane_bundles.h (kerrtrace shim) -- ImagePlaneBundles<T>: five rays (centre, E, W, N, S) per image-plane
// point for local Jacobians, with the reference's public members and accessors
// (src/raytracer/imageplane_bundles.h:44-197).  Initialisation runs on the GPU.
#ifndef IMAGEPLANE_BUNDLES_H_
#define IMAGEPLANE_BUNDLES_H_

#include <cmath>
#include <iostream>

#include "imageplane.h"

template <typename T>
class ImagePlaneBundles : public Raytracer<T>
{
public:
    static constexpr int RAYS_PER_BUNDLE = 5;

    int Nx, Ny;        // grid dimensions (fencepost: Nx+1 points along each axis)
    T eps_x, eps_y;    // satellite offsets (rg)

private:
    T D, incl, phi0;
    T m_x0, m_dx, m_y0, m_dy;

public:
    ImagePlaneBundles(T dist, T inc_deg, T x0, T xmax, T dx, T y0, T ymax, T dy, T spin, T phi, T precision = PRECISION,
                      T eps_frac = 0.01)
        : Raytracer<T>(kerrtrace_shim::plane_count<T>(kerrtrace_shim::describe_plane<T>(dist, inc_deg, x0, xmax, dx, y0, ymax, dy, phi, true, eps_frac)),
                       -1 * spin, precision),
          eps_x(eps_frac * dx), eps_y(eps_frac * dy), D(dist), incl(inc_deg * M_PI / 180.0), phi0(phi), m_x0(x0), m_dx(dx), m_y0(y0), m_dy(dy)
    {
        const kt_imageplane p = kerrtrace_shim::describe_plane<T>(dist, inc_deg, x0, xmax, dx, y0, ymax, dy, phi, true, eps_frac);
        kerrtrace_shim::plane_count<T>(p, &Nx, &Ny);
        if (!kerrtrace_shim::quiet())
            std::cout << "Setting up image plane bundles with (" << Nx << 'x' << Ny << ") bundle centres, " << RAYS_PER_BUNDLE
                      << " rays each = " << Nx * Ny * RAYS_PER_BUNDLE << " total rays" << std::endl
                      << "  bundle eps_x=" << eps_x << " eps_y=" << eps_y << "  (eps_frac=" << eps_frac << ")" << std::endl;
        kerrtrace_shim::check(kt_imageplane_init(this->engine(), this->tracer(), &p, this->as_kt(), this->nRays), "init bundles");
    }

    inline int centre_ray(int ix, int iy) const { return (ix * Ny + iy) * RAYS_PER_BUNDLE + 0; }
    inline int east_ray(int ix, int iy) const { return (ix * Ny + iy) * RAYS_PER_BUNDLE + 1; }
    inline int west_ray(int ix, int iy) const { return (ix * Ny + iy) * RAYS_PER_BUNDLE + 2; }
    inline int north_ray(int ix, int iy) const { return (ix * Ny + iy) * RAYS_PER_BUNDLE + 3; }
    inline int south_ray(int ix, int iy) const { return (ix * Ny + iy) * RAYS_PER_BUNDLE + 4; }

    void redshift_start() { Raytracer<T>::redshift_start(0, true); }
    using Raytracer<T>::redshift;
};

#endif /* IMAGEPLANE_BUNDLES_H_ */
