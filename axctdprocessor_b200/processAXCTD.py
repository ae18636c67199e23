"""Drop-in mirror of the reference command line (reference processAXCTD.py:47-183):
same flags, same settings dict, same output-file format.

Two extra, optional flags that the reference does not have:
  --wired     make -s/-e/-a/-b/-p/-t/-l/-u act as documented (as shipped they are
              inert or crash); default is faithful behaviour
  --device N  CUDA device index
"""
from __future__ import annotations

import argparse
import logging
import os

from . import AXCTDprocessor


def main(argv=None):
    parser = argparse.ArgumentParser(description='Demodulate an audio file to text')
    parser.add_argument('-i', '--input', default='ERROR_NO_FILE_SPECIFIED', help='Input WAV filename')
    parser.add_argument('-o', '--output', default='output.txt', help='Output filename')
    parser.add_argument('-s', '--starttime', default='0', help='AXCTD start time in WAV file')
    parser.add_argument('-e', '--endtime', default='-1', help='AXCTD end time in WAV file')
    parser.add_argument('-a', '--autodetect-start', default='30', help='Point at which autodetect algorithm starts scanning for profile transmission start')
    parser.add_argument('-b', '--autodetect-end', default='-1', help='Point at which autodetect algorithm stops scanning for profile transmission start')
    parser.add_argument('-p', '--sig-threshold-400', default='2', help='Threshold for normalized 400 Hz signal level to detect profile transmission')
    parser.add_argument('-t', '--sig-threshold-7500', default='1.5', help='Threshold for normalized 7500 Hz signal level to detect profile transmission')
    parser.add_argument('-d', '--dead-freq', default='3000', help='"Dead" (quiet) frequency used to calculate normalized signal levels (Hz)')
    parser.add_argument('-l', '--pointsperloop', default='100000', help='Number of PCM audio data points processed per iteration')
    parser.add_argument('-m', '--mark-freq', default='400', help='Mark (bit 1) frequency (Hz)')
    parser.add_argument('-n', '--space-freq', default='800', help='Space (bit 0) frequency (Hz)')
    parser.add_argument('-u', '--use-bandpass', action='store_true', help='Apply this flag to use a bandpass filter (100 Hz to 1200 Hz) rather than a 1200 Hz lowpass filter before demodulation')
    parser.add_argument('--wired', action='store_true', help='make the documented flags act (not reference behaviour)')
    parser.add_argument('--device', type=int, default=0, help='CUDA device index')
    args = parser.parse_args(argv)

    if args.input == 'ERROR_NO_FILE_SPECIFIED':
        print("[!] Error- no input WAV file specified! Terminating")
        exit()
    elif not os.path.exists(args.input):
        print("[!] Specified input file does not exist! Terminating")
        exit()

    timerange = [parse_times(args.starttime), parse_times(args.endtime)]      # processAXCTD.py:80-84
    if timerange[1] <= 0:
        timerange[1] = -1
    triggerrange = [parse_times(args.autodetect_start), parse_times(args.autodetect_end)]   # :87-91
    if triggerrange[1] <= 0:
        triggerrange[1] = -1

    settings = {'triggerrange': triggerrange,                                # :93-99
                'minR400': float(args.sig_threshold_400),
                'mindR7500': float(args.sig_threshold_7500),
                'deadfreq': float(args.dead_freq),
                'pointsperloop': int(args.pointsperloop),
                'mark_space_freqs': [float(args.mark_freq), float(args.space_freq)],
                'use_bandpass': args.use_bandpass}
    return processAXCTD(args.input, args.output, timerange, settings,
                        mode="wired" if args.wired else "faithful", device=args.device)


def parse_times(time_string):
    """processAXCTD.py:106-121."""
    try:
        if ":" in time_string:
            t = 0
            for i, val in enumerate(reversed(time_string.split(":"))):
                if i <= 2:
                    t += int(val) * 60 ** i
                else:
                    logging.info("[!] Warning- ignoring all end time information past the hours place (HH:MM:SS)")
        else:
            t = int(time_string)
        return t
    except ValueError:
        logging.info("[!] Unable to interpret specified start time- defaulting to 00:00")
        return -2


def wired_settings(settings, f_s):
    """CLI key names -> the processor's internal names (what the flags document)."""
    return {'minr400': settings['minR400'], 'mindr7500': settings['mindR7500'], 'deadfreq': settings['deadfreq'],
            'refreshrate': settings['pointsperloop'] / f_s, 'mark_space_freqs': settings['mark_space_freqs'],
            'usebandpass': settings['use_bandpass']}


def profile_lines(ap, wavfile, timerange, settings, defaults=None):
    """Lines of the output file in order, processAXCTD.py:146-183.  ``defaults``
    supplies the '(default)' coefficient lines; without it an incomplete header
    raises KeyError('zcoeff_default') at the same point as the reference
    (:161-167), after the lines before it were produced."""
    minR400 = settings['minR400']
    mindR7500 = settings['mindR7500']
    deadfreq = settings['deadfreq']
    pointsperloop = settings['pointsperloop']
    triggerrange = settings['triggerrange']
    yield f"AXCTD profile for {wavfile}\n"
    fs = ap.f_s
    yield f'Sampling frequency (fs): {fs} Hz\n'
    yield f'Audio file length: {ap.numpoints/fs} sec\n'
    yield f'400 Hz pulse start: {ap.firstpulse400/fs} sec\n'
    yield f'7500 Hz tone start: {ap.profstartind/fs} sec\n'
    yield "\nAXCTD header information:\n"
    for desc, ckey in zip(['Probe Code', 'Maximum Depth (m)', 'Probe Serial'], ['probe_code', 'max_depth', 'serial_no']):
        yield f"{desc}: {ap.metadata[ckey]}\n"
    yield "Conversion equations:\n"
    md = dict(ap.metadata)
    if defaults:
        md.update(defaults)
    for coeff, desc, symb in zip(['z', 't', 'c'], ['Depth', 'Temperature', 'Conductivity'], ['t', 'T', 'C']):
        if sum(ap.metadata[coeff + 'coeff_valid']) == 4:
            cfield = coeff + 'coeff'
            defaultstatus = ''
        else:
            cfield = coeff + 'coeff_default'
            defaultstatus = '(default)'
        cureqn = ' + '.join([f'{val}*{symb}^{i}' for i, val in enumerate(md[cfield])])
        yield f'{desc}: {cureqn} {defaultstatus}\n'
    yield '\nProcessor Settings:\n'
    yield f'Time Range: {timerange[0]} sec to {timerange[1] if timerange[1] >= 0 else "N/A"} sec\n'
    yield f'Min. 400 Hz power ratio: {minR400}\n'
    yield f'Min. 7500 Hz power ratio: {mindR7500}\n'
    yield f'Dead frequency: {deadfreq}\n'
    yield f'Points per loop: {pointsperloop}\n'
    yield f'Trigger range: {triggerrange[0]} sec to {triggerrange[1] if triggerrange[1] >= 0 else "N/A"} sec\n'
    yield '\nAXCTD Profile:\n'
    yield 'Time (s), Hex Frame, Depth (m), Temperature (C), Conductivity (mS/cm), Salinity (PSU)\n'
    for (t, hf, z, T, C, S) in zip(ap.time, ap.hexframes, ap.depth, ap.temperature, ap.conductivity, ap.salinity):
        yield f"{t:8.2f},  {hf},{z:10.2f},{T:16.2f},{C:21.2f},{S:15.2f}\n"


def format_profile(ap, wavfile, timerange, settings, defaults=None):
    return "".join(profile_lines(ap, wavfile, timerange, settings, defaults))


def processAXCTD(wavfile, outfile, timerange, settings, mode="faithful", device=0, engine=None):
    """processAXCTD.py:126-183."""
    for key in ('minR400', 'mindR7500', 'deadfreq', 'pointsperloop', 'triggerrange', 'mark_space_freqs', 'use_bandpass'):
        settings[key]                                                        # KeyError like :128-134
    print("Processing profile")
    if mode == "wired":
        probe_fs = AXCTDprocessor.read_wav_pcm16(wavfile)[0]
        ap = AXCTDprocessor.AXCTD_Processor(wavfile, timerange=timerange, user_settings=wired_settings(settings, probe_fs),
                                            mode="wired", device=device, engine=engine)
        ap.triggerrange = list(settings['triggerrange'])
        defaults = {c + 'coeff_default': ap.settings[c + 'coeff_axctd'] for c in 'ztc'}
    else:
        ap = AXCTDprocessor.AXCTD_Processor(wavfile, timerange=timerange, user_settings=settings, device=device, engine=engine)
        defaults = None
    ap.run()
    print("Profile processing complete- writing output files")
    with open(outfile, 'w') as f:
        # line by line, so that a KeyError leaves the same partial file the reference leaves
        for line in profile_lines(ap, wavfile, timerange, settings, defaults):
            f.write(line)


if __name__ == "__main__":
    main()
