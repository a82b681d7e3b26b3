#pragma once
extern "C" { double gsl_sf_hyperg_2F1(double a, double b, double c, double x); }
