// pointsource.h (kerrtrace shim) -- PointSource<T> with the reference's constructor and methods
// (src/raytracer/pointsource.h:24-32, pointsource.cpp:11-86).  The ray grid is initialised on the GPU.
#ifndef POINTSOURCE_H_
#define POINTSOURCE_H_

#include "raytracer.h"
#include "ray_destination.h"

template <typename T>
class PointSource : public Raytracer<T>
{
private:
    T energy;
    T velocity;
    int n_cosalpha;
    int n_beta;

    static kt_pointsource describe(T* pos, T V, T dcosalpha, T dbeta, T cosalpha0, T cosalphamax, T beta0, T betamax, T E)
    {
        kt_pointsource s = {};
        for (int i = 0; i < 4; i++) s.pos[i] = pos[i];
        s.V = V; s.E = E;
        s.dcosalpha = dcosalpha; s.dbeta = dbeta;
        s.cosalpha0 = cosalpha0; s.cosalphamax = cosalphamax;
        s.beta0 = beta0; s.betamax = betamax;
        return s;
    }
    static int count(T* pos, T V, T dcosalpha, T dbeta, T cosalpha0, T cosalphamax, T beta0, T betamax, T E)
    {
        const kt_pointsource s = describe(pos, V, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, E);
        return static_cast<int>(kt_pointsource_count(&s, nullptr, nullptr));
    }

public:
    PointSource(T* pos, T V, T spin, T tol, T dcosalpha, T dbeta, T cosalpha0 = -0.999999, T cosalphamax = 0.995,
                T beta0 = -0.995 * M_PI, T betamax = M_PI, T E = 1)
        : Raytracer<T>(count(pos, V, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, E), spin, tol), energy(E), velocity(V)
    {
        if (!kerrtrace_shim::quiet()) cout << "Setting up point source with " << Raytracer<T>::nRays << " rays" << endl;
        init_pointsource(pos, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax);
    }

    void init_pointsource(T* pos, T dcosalpha, T dbeta, T cosalpha0 = -0.999999, T cosalphamax = 0.995, T beta0 = -0.995 * M_PI,
                          T betamax = M_PI)
    {
        const kt_pointsource s = describe(pos, velocity, dcosalpha, dbeta, cosalpha0, cosalphamax, beta0, betamax, energy);
        int32_t na = 0, nb = 0;
        kt_pointsource_count(&s, &na, &nb);
        n_cosalpha = na;
        n_beta = nb;
        kerrtrace_shim::check(kt_pointsource_init(this->engine(), this->tracer(), &s, this->as_kt(), this->nRays), "init_pointsource");
    }

    void redshift_start() { Raytracer<T>::redshift_start(velocity); }
    void redshift(T V) { Raytracer<T>::redshift(V); }
    using Raytracer<T>::redshift;
};

#endif /* POINTSOURCE_H_ */
