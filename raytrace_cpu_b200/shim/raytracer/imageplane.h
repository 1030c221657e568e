// imageplane.h (kerrtrace shim) -- ImagePlane<T> with the reference's constructor and methods
// (src/raytracer/imageplane.h:27-60, imageplane.cpp:11-140).  Rays are traced backwards from the observer, so the
// base class stores the NEGATED spin exactly like the reference (imageplane.cpp:12); initialisation runs on the GPU.
#ifndef IMAGEPLANE_H_
#define IMAGEPLANE_H_

#include "raytracer.h"
#include "ray_destination.h"

namespace kerrtrace_shim {
template <typename T>
kt_imageplane describe_plane(T dist, T inc, T x0, T xmax, T dx, T y0, T ymax, T dy, T phi, bool bundles, T eps_frac)
{
    kt_imageplane p = {};
    p.dist = dist; p.incl_deg = inc;
    p.x0 = x0; p.xmax = xmax; p.dx = dx;
    p.y0 = y0; p.ymax = ymax; p.dy = dy;
    p.phi0 = phi;
    p.bundles = bundles ? 1 : 0;
    p.eps_frac = eps_frac;
    return p;
}
template <typename T>
int plane_count(const kt_imageplane& p, int* nx = nullptr, int* ny = nullptr)
{
    int32_t a = 0, b = 0;
    const int n = static_cast<int>(kt_imageplane_count(&p, &a, &b));
    if (nx) *nx = a;
    if (ny) *ny = b;
    return n;
}
}  // namespace kerrtrace_shim

template <typename T>
class ImagePlane : public Raytracer<T>
{
private:
    T D, incl, phi0;
    T m_x0, m_xmax, m_dx;
    T m_y0, m_ymax, m_dy;
    int Nx, Ny;

public:
    ImagePlane(T dist, T inc, T x0, T xmax, T dx, T y0, T ymax, T dy, T spin, T phi, T precision = PRECISION)
        : Raytracer<T>(kerrtrace_shim::plane_count<T>(kerrtrace_shim::describe_plane<T>(dist, inc, x0, xmax, dx, y0, ymax, dy, phi, false, 0)),
                       -1 * spin, precision),
          D(dist), incl(inc), phi0(phi), m_x0(x0), m_xmax(xmax), m_dx(dx), m_y0(y0), m_ymax(ymax), m_dy(dy)
    {
        const kt_imageplane p = kerrtrace_shim::describe_plane<T>(dist, inc, x0, xmax, dx, y0, ymax, dy, phi, false, 0);
        kerrtrace_shim::plane_count<T>(p, &Nx, &Ny);
        if (!kerrtrace_shim::quiet()) cout << "Setting up image plane with (" << Nx << 'x' << Ny << ") rays" << endl;
        kerrtrace_shim::check(kt_imageplane_init(this->engine(), this->tracer(), &p, this->as_kt(), this->nRays), "init_image_plane");
    }

    void redshift_start() { Raytracer<T>::redshift_start(0, true); }
    void redshift(bool projradius) { Raytracer<T>::redshift(-1, true, projradius); }
    using Raytracer<T>::redshift;

    inline int get_x_index(int ix) { return static_cast<int>(ix / Ny); }
    inline int get_y_index(int ix) { return ix % Ny; }
    inline T ray_x(int ix) { return m_x0 + get_x_index(ix) * m_dx; }
    inline T ray_y(int ix) { return m_y0 + get_y_index(ix) * m_dy; }
};

#endif /* IMAGEPLANE_H_ */
