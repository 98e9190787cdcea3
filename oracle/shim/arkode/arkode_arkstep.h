/* TEST INFRASTRUCTURE: the ARKStep entry points the reference's euler3D_main.cpp calls
 * (euler3D_main.cpp:191-449), implemented in oracle/shim/shim_arkstep.cpp on top of the explicit
 * Runge-Kutta loop of this repository (host/erk_stepper.hpp).  It is NOT ARKODE: it exists so that
 * the reference's own main program, problem files and io.cpp run unmodified on the CPU next to the
 * native driver, with the same time stepper on both sides. */
#ifndef EULERB200_SHIM_ARKSTEP_H
#define EULERB200_SHIM_ARKSTEP_H
#include "shim_core.h"
#include <cstdio>
#define ARK_NORMAL 1
#define ARK_SUCCESS 0
typedef int (*ARKRhsFn)(realtype t, N_Vector y, N_Vector ydot, void* user_data);
typedef int (*ARKExpStabFn)(N_Vector y, realtype t, realtype* hstab, void* user_data);
void* ARKStepCreate(ARKRhsFn fe, ARKRhsFn fi, realtype t0, N_Vector y0, SUNContext ctx);
void ARKStepFree(void** arkode_mem);
int ARKStepSetUserData(void*, void* user_data);
int ARKStepSetDiagnostics(void*, FILE*);
int ARKStepSetOrder(void*, int order);
int ARKStepSetTableNum(void*, ARKODE_DIRKTableID itable, ARKODE_ERKTableID etable);
int ARKStepSetDenseOrder(void*, int);
int ARKStepSetSafetyFactor(void*, realtype);
int ARKStepSetErrorBias(void*, realtype);
int ARKStepSetMaxGrowth(void*, realtype);
int ARKStepSetAdaptivityMethod(void*, int imethod, int idefault, int pq, realtype* params);
int ARKStepSetMaxFirstGrowth(void*, realtype);
int ARKStepSetMaxEFailGrowth(void*, realtype);
int ARKStepSetInitStep(void*, realtype);
int ARKStepSetMinStep(void*, realtype);
int ARKStepSetMaxStep(void*, realtype);
int ARKStepSetMaxErrTestFails(void*, int);
int ARKStepSetMaxHnilWarns(void*, int);
int ARKStepSetStabilityFn(void*, ARKExpStabFn, void* data);
int ARKStepSetFixedStep(void*, realtype h);
int ARKStepSetMaxNumSteps(void*, long int);
int ARKStepSStolerances(void*, realtype rtol, realtype atol);
int ARKStepSetStopTime(void*, realtype tstop);
int ARKStepEvolve(void*, realtype tout, N_Vector yout, realtype* tret, int itask);
int ARKStepGetCurrentStep(void*, realtype* hcur);
int ARKStepGetNumSteps(void*, long int*);
int ARKStepGetNumStepAttempts(void*, long int*);
int ARKStepGetNumRhsEvals(void*, long int* nfe, long int* nfi);
int ARKStepGetNumErrTestFails(void*, long int*);
#endif
